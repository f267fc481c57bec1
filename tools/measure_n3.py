#!/usr/bin/env python3
"""Row N3 measurements (one JSON object): a 1M BIP-340 batch verified one by one (sv_verify_host) and by random linear
combination (sv_verify_schnorr_batch_host) with no, sparse (0.01 %) and heavy (10 %) damage; verdicts compared with the
reference on a sample.  Host buffers, wall clock around the synchronous calls, inputs pinned."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightning_b200 as L  # noqa: E402
from tests import util  # noqa: E402


def main():
    n = 1_000_000
    eng = L.SigVerifier(0)
    eng.set_profiling(True)
    ref = util.load_ref()
    p8 = ctypes.POINTER(ctypes.c_uint8)
    msg, key, sig = np.zeros((n, 32), np.uint8), np.zeros((n, 32), np.uint8), np.zeros((n, 64), np.uint8)
    ref.ref_make_schnorr_batch(ctypes.c_uint64(99), ctypes.c_size_t(n), msg.ctypes.data_as(p8), key.ctypes.data_as(p8), sig.ctypes.data_as(p8), 16)
    damaged = sig.copy()  # ref_make_schnorr_batch corrupts every 10th item in place: keep that as the heavy case ...
    clean_idx = np.arange(n) % 10 != 0
    cm, ck, cs = (np.ascontiguousarray(a[clean_idx]) for a in (msg, key, sig))  # ... and the untouched 900k as the clean batch
    out = {"n_clean": int(cm.shape[0])}

    def pin(a):
        b = eng.host_alloc(a.nbytes)
        b[:] = a.reshape(-1)
        return b.reshape(a.shape)
    msg, key, sig, damaged, cm, ck, cs = (pin(a) for a in (msg, key, sig, damaged, cm, ck, cs))

    def timed(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        dt = (time.perf_counter() - t0) / reps
        timed.dev = [round(x, 3) for x in eng.last_timing()]  # device time of the last call: (scalar side / preparation, curve side / bucket sums)
        return r, dt
    m = cm.shape[0]
    v, dt = timed(lambda: eng.verify(2, cm, ck, cs))
    assert v.all()
    out["one_by_one"] = {"verifies_per_s": m / dt, "ms": dt * 1e3, "device_ms_prep_main": timed.dev}
    (v, gt, gf), dt = timed(lambda: eng.verify_schnorr_batch(cm, ck, cs))
    assert v.all() and gf == 0
    out["batch_all_valid"] = {"verifies_per_s": m / dt, "ms": dt * 1e3, "groups": gt, "groups_failed": gf, "device_ms_prep_buckets": timed.dev}
    sm = pin(np.array(cm))
    bad = np.arange(0, m, 10_000)
    sm[bad, 3] ^= 1
    (v, gt, gf), dt = timed(lambda: eng.verify_schnorr_batch(sm, ck, cs))
    assert (v == 0).sum() == bad.size and not v[bad].any()
    out["batch_0.01pct_bad"] = {"verifies_per_s": m / dt, "ms": dt * 1e3, "groups": gt, "groups_failed": gf}
    (v, gt, gf), dt = timed(lambda: eng.verify_schnorr_batch(msg, key, damaged))
    want = util.ref_verify(ref, 2, msg[:50000], key[:50000], damaged[:50000], threads=16)
    assert np.array_equal(v[:50000], want)
    out["batch_10pct_bad"] = {"verifies_per_s": n / dt, "ms": dt * 1e3, "groups": gt, "groups_failed": gf, "valid_fraction": float(v.mean())}
    v1, dt = timed(lambda: eng.verify(2, msg, key, damaged))
    assert np.array_equal(v1, v)
    out["one_by_one_10pct_bad"] = {"verifies_per_s": n / dt, "ms": dt * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
