#!/usr/bin/env python3
"""Build variants of the engine (different -D flags) into variants/<name>/libcln_sigverify.so and, on a GPU box, measure each
with the device-resident 1M ECDSA33 step of bench.py's workload (two launch streams, CUDA events).

  python tools/variants.py build            (here: nvcc cross-compiles without a GPU)
  python tools/variants.py measure          (under gpurun) -> JSON lines
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "variants")
VARIANTS = {
    "base": [],
    "alu_folds": ["-DSV_ALU_FOLDS"],
    "reduce_noacc": ["-DSV_REDUCE_NOACC"],
    "alu_folds+noacc": ["-DSV_ALU_FOLDS", "-DSV_REDUCE_NOACC"],
    # the 8-bit GLV comb staged in shared memory by one bulk (TMA) copy: 1 CTA of 512 threads per SM (136 KiB of shared memory)
    "comb_smem": ["-DSV_COMB_SMEM", "-DSV_MAIN_BLOCK=512", "-DSV_MAIN_MINB=1"],
    "sync_w1": ["-USV_SYNC_WINDOWS", "-DSV_SYNC_WINDOWS=1"],
    "sync_w4": ["-USV_SYNC_WINDOWS", "-DSV_SYNC_WINDOWS=4"],
}


def build():
    from lightning_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    dropin_o = os.path.join(b.CSRC, "cln_dropin.o")
    for name, flags in VARIANTS.items():
        d = os.path.join(VDIR, name)
        os.makedirs(d, exist_ok=True)
        base = [f for f in b.NVCC_FLAGS if not (name.startswith("sync_w") and f.startswith("-DSV_SYNC_WINDOWS"))]
        cmd = [nvcc] + base + [f for f in flags if not f.startswith("-U")] + ["-o", os.path.join(d, "libcln_sigverify.so"),
                                                                             os.path.join(b.CSRC, "engine.cu"), os.path.join(b.CSRC, "batch.o"), dropin_o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(name, "ok" if r.returncode == 0 else "FAILED\n" + r.stderr[-2000:])


def measure():
    names = sys.argv[2:] or list(VARIANTS)
    for name in names:
        lib = os.path.join(VDIR, name, "libcln_sigverify.so")
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, SV_LIB=lib, SV_BENCH_QUICK="1")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "30", "--warmup", "3"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(json.dumps({"variant": name, "flags": VARIANTS[name], "value": d["value"], "kernel_ms": d["roofline"]["kernel_ms"],
                              "prep_ms": d["roofline"]["prep_kernel_ms"], "regs": d["engine"]["main_regs"], "checks": d.get("failed_checks")}))
        except Exception as ex:  # noqa: BLE001
            print(json.dumps({"variant": name, "error": repr(ex), "stderr": r.stderr[-500:]}))


if __name__ == "__main__":
    {"build": build, "measure": measure}[sys.argv[1]]()
