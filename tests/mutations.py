"""Structured + random mutations of valid triples for differential testing (engine / host build vs the reference)."""
import numpy as np

from tests import util

N, P = util.N_ORDER, util.P_FIELD


def _be(v):
    return np.frombuffer(int(v % (1 << 256)).to_bytes(32, "big"), dtype=np.uint8)


def mutate(w, seed):
    """In place: every item receives one mutation chosen from ~40 classes (boundary values of r, s, x, the message,
    swapped fields, random bit flips at random positions ...).  Returns the class index per item."""
    rng = np.random.default_rng(seed)
    n = w["msg"].shape[0]
    specials = [0, 1, 2, N - 1, N, N + 1, (N - 1) // 2, (N + 1) // 2, (N + 1) // 2 + 1, P - N - 1, P - N, P - N + 1, P - 1, P, P + 1,
                2**256 - 1, 2**255, 2**128, 2**32 + 977]
    cls = np.zeros(n, np.int32)
    for i in range(n):
        c = int(rng.integers(0, 40))
        cls[i] = c
        if c < 8:  # random single bit flip somewhere
            field = [("msg", 32), ("sig", 64), ("ssig", 64), ("pub33", 33), ("pubxy", 64), ("xonly", 32)][c % 6]
            pos, bit = int(rng.integers(0, field[1])), int(rng.integers(0, 8))
            w[field[0]][i, pos] ^= 1 << bit
        elif c < 14:  # r := special
            v = specials[int(rng.integers(0, len(specials)))]
            w["sig"][i, :32] = _be(v); w["ssig"][i, :32] = _be(v)
        elif c < 20:  # s := special
            v = specials[int(rng.integers(0, len(specials)))]
            w["sig"][i, 32:] = _be(v); w["ssig"][i, 32:] = _be(v)
        elif c < 24:  # key x := special
            v = specials[int(rng.integers(0, len(specials)))]
            w["pub33"][i, 1:] = _be(v); w["pubxy"][i, :32] = _be(v); w["xonly"][i] = _be(v)
        elif c < 26:  # message := special (valid for the hash to be >= n: it is reduced, never rejected)
            v = specials[int(rng.integers(0, len(specials)))]
            w["msg"][i] = _be(v)
        elif c < 28:  # prefix byte
            w["pub33"][i, 0] = int(rng.integers(0, 256))
        elif c < 30:  # y := p - y  (other point of the same x; xy form must now fail, 33-byte form flips parity)
            y = int.from_bytes(bytes(w["pubxy"][i, 32:]), "big")
            w["pubxy"][i, 32:] = _be((P - y) % P)
            w["pub33"][i, 0] ^= 1
        elif c < 32:  # s := n - s (high-S / negated)
            s = int.from_bytes(bytes(w["sig"][i, 32:]), "big")
            w["sig"][i, 32:] = _be((N - s) % N)
            s = int.from_bytes(bytes(w["ssig"][i, 32:]), "big")
            w["ssig"][i, 32:] = _be((N - s) % N)
        elif c < 34:  # swap r and s
            w["sig"][i] = np.concatenate([w["sig"][i, 32:], w["sig"][i, :32]])
            w["ssig"][i] = np.concatenate([w["ssig"][i, 32:], w["ssig"][i, :32]])
        elif c < 36:  # all-zero / all-ones field
            f = ["msg", "sig", "pub33", "xonly"][int(rng.integers(0, 4))]
            w[f][i] = 0 if c == 34 else 255
            if f == "sig":
                w["ssig"][i] = w["sig"][i]
        # 36..39: leave valid
    return cls
