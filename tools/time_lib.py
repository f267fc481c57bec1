"""Time sv_verify_device of an arbitrary build of the library (only the core symbols are used)."""
import ctypes, sys, os
import torch
lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
vp, sz, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
lib.sv_create.argtypes = [ctypes.POINTER(vp), i]
lib.sv_synth_device.argtypes = [vp, i, ctypes.c_uint64, sz, vp, vp, vp, vp]
lib.sv_verify_device.argtypes = [vp, i, vp, vp, vp, sz, vp, vp, vp]
lib.sv_sync.argtypes = [vp, vp]
lib.sv_get_stream.argtypes = [vp]; lib.sv_get_stream.restype = vp
ctx = vp()
assert lib.sv_create(ctypes.byref(ctx), 0) == 0
msg = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); key = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
sig = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); ver = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
assert lib.sv_synth_device(ctx, 0, 42, n, msg.data_ptr(), key.data_ptr(), sig.data_ptr(), None) == 0
lib.sv_sync(ctx, None)
ext = torch.cuda.ExternalStream(lib.sv_get_stream(ctx))
best = 1e9
for rep in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    assert lib.sv_verify_device(ctx, 0, msg.data_ptr(), key.data_ptr(), sig.data_ptr(), n, ver.data_ptr(), None, None) == 0
    e1.record(ext)
    lib.sv_sync(ctx, None)
    if rep: best = min(best, e0.elapsed_time(e1))
print(sys.argv[1], "n", n, "%.2f ms  %.2f Mverify/s  valid %d" % (best, n / best / 1e3, int(ver.sum().item())))
