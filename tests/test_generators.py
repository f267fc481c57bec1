"""CPU-only: the code generators reproduce what is committed, and the generated PTX programs are right.

`tools/gen_mul.py` writes the inline-PTX bodies of the 256x256 product, the square and the scalar fold
(`lightning_b200/csrc/u256_gen.cuh`); `--check` simulates those programs (carry-flag semantics of mul/mad/madc/add/addc)
against Python integers.  `tools/gen_wire.py` writes the verifier daemon's wire codec from its CSV."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable] + list(args), cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_generated_ptx_programs_match_bigint_arithmetic():
    r = run("tools/gen_mul.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "match" in r.stdout


def test_committed_generated_header_is_what_the_generator_emits():
    r = run("tools/gen_mul.py")
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(ROOT, "lightning_b200", "csrc", "u256_gen.cuh")).read()
