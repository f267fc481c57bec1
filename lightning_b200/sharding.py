"""Host-side logic of the multi-GPU path: verifications are independent, so a batch is cut
into contiguous per-rank slices (no data-path collective) and the only exchange is the gather of the
1-bit-per-verification verdict bitmap."""
import numpy as np


def shard_range(n_total, rank, world):
    """[lo, hi) of rank's contiguous slice; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_shard(n_total, world):
    return -(-n_total // world)


def bitmap_words(n):
    return (n + 31) // 32


def pack_bitmap(verdict_bytes):
    """uint8 0/1 vector -> little-endian 32-bit words, bit i%32 of word i//32 (the layout k_pack_bitmap writes)."""
    v = np.asarray(verdict_bytes, dtype=np.uint8)
    pad = (-v.size) % 32
    v = np.concatenate([v, np.zeros(pad, np.uint8)]).reshape(-1, 32).astype(np.uint32)
    return (v << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint64).astype(np.uint32)


def unpack_bitmap(words, n):
    w = np.asarray(words, dtype=np.uint32)
    return ((w[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:n].astype(np.uint8)


def unpack_gathered(words, n_total, world):
    """Concatenate the per-rank bitmaps of an all_gather (each padded to bitmap_words(max_shard)) into the
    job-wide verdict vector."""
    per = bitmap_words(max_shard(n_total, world))
    out = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out.append(unpack_bitmap(words[r * per:(r + 1) * per], hi - lo))
    return np.concatenate(out)
