"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU path — contiguous sharding of
the triple array and the gather of per-rank verdict bitmaps into the job-wide bitmap (bench.py, §8e)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from lightning_b200 import sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 100_003
    lo, hi = sharding.shard_range(n_total, rank, world)
    # every rank derives the same global verdict vector; its shard plays the role of the kernel output
    g = np.random.default_rng(5).integers(0, 2, size=n_total, dtype=np.uint8)
    local_bits = torch.from_numpy(sharding.pack_bitmap(g[lo:hi]).view(np.int32).copy())
    words = sharding.bitmap_words(sharding.max_shard(n_total, world))
    padded = torch.zeros(words, dtype=torch.int32)
    padded[: local_bits.numel()] = local_bits
    out = torch.zeros(world * words, dtype=torch.int32)
    dist.all_gather_into_tensor(out, padded)
    full = sharding.unpack_gathered(out.numpy().view(np.uint32), n_total, world)
    assert np.array_equal(full, g), "gathered bitmap differs from the global verdict vector"
    assert sum(sharding.shard_range(n_total, r, world)[1] - sharding.shard_range(n_total, r, world)[0] for r in range(world)) == n_total
    dist.barrier()
    if rank == 0:
        print("OK")
""")


def test_shard_and_gather_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
