"""Integer-pipe probes + a quick device-resident throughput number (development script)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lightning_b200 as L

eng = L.SigVerifier(0)
print("info", eng.info())
names = ["imad_wide MAC/s", "cmad4 MAC/s", "fe_mul/s", "fe_sqr/s", "chain8 MAC/s", "carry_save MAC/s", "imad32 instr/s", "addc adds/s", "dfma /s"]
if "--noprobe" not in sys.argv:
    for mode, name in enumerate(names):
        v = eng.probe(mode)
        print("probe %-18s %.4g   (%.1f /clk/SM @1.965GHz)" % (name, v, v / 148 / 1.965e9))
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 20
ext = torch.cuda.ExternalStream(eng.stream_handle())
for kind, ks in [(0, 33), (1, 64), (2, 32)]:
    msg = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    key = torch.empty(n * ks, dtype=torch.uint8, device="cuda")
    sig = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ver = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    eng.synth_device(kind, 42, n, msg.data_ptr(), key.data_ptr(), sig.data_ptr())
    eng.sync()
    print("synth kind", kind, "%.1f ms" % ((time.time() - t0) * 1e3))
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        eng.verify_device(kind, msg.data_ptr(), key.data_ptr(), sig.data_ptr(), n, ver.data_ptr())
        e1.record(ext)
        eng.sync()
        ms = e0.elapsed_time(e1)
        print("kind", kind, "n", n, "%.2f ms" % ms, "%.3f Mverify/s" % (n / ms / 1e3), "valid", int(ver.sum().item()))
