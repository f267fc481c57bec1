"""First-contact GPU script: integer-pipe probes + a quick device-resident throughput number."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lightning_b200 as L

eng = L.SigVerifier(0)
print("info", eng.info())
for mode, name in [(0, "imad_wide_indep MAC/s"), (1, "imad_wide_carry MAC/s"), (2, "fe_mul/s"), (3, "fe_sqr/s")]:
    print(name, "%.4g" % eng.probe(mode))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for kind, ks in [(0, 33), (1, 64), (2, 32)]:
    msg = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    key = torch.empty(n * ks, dtype=torch.uint8, device="cuda")
    sig = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ver = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    eng.synth_device(kind, 42, n, msg.data_ptr(), key.data_ptr(), sig.data_ptr(), st)
    torch.cuda.synchronize()
    print("synth kind", kind, "%.1f ms" % ((time.time() - t0) * 1e3))
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.verify_device(kind, msg.data_ptr(), key.data_ptr(), sig.data_ptr(), n, ver.data_ptr(), 0, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("kind", kind, "n", n, "%.2f ms" % ms, "%.3f Mverify/s" % (n / ms / 1e3), "valid", int(ver.sum().item()))
