"""Adversarial-but-VALID signatures with prescribed (u1, u2), built with plain Python integers.

An attacker who knows the private key controls both scalars of the verification equation
R = u1*G + u2*Q (choose R, solve for s and the message hash).  These inputs steer the fixed-window
ladder and the comb through their exceptional branches (accumulator == +-table point, infinity,
zero digits, top-window carries, GLV lattice vectors ...), where a verifier that mishandles a case
would wrongly reject a signature the reference accepts."""
import numpy as np

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72


def inv(x, m=P):
    return pow(x, m - 2, m)


def add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        l = 3 * a[0] * a[0] * inv(2 * a[1]) % P
    else:
        l = (b[1] - a[1]) * inv(b[0] - a[0]) % P
    x = (l * l - a[0] - b[0]) % P
    return (x, (l * (a[0] - x) - a[1]) % P)


def mul(k, pt):
    k %= N
    r = None
    while k:
        if k & 1:
            r = add(r, pt)
        pt = add(pt, pt)
        k >>= 1
    return r


G = (GX, GY)


def craft(d, u1, u2):
    """Return (msg32, pub33, pubxy, sig64) of a signature that verifies with scalars (+-u1, +-u2), or None."""
    u1 %= N
    u2 %= N
    if u2 == 0 or d % N == 0:
        return None
    Q = mul(d, G)
    R = mul((u1 + u2 * d) % N, G)
    if R is None:
        return None
    r = R[0] % N
    if r == 0:
        return None
    s = r * inv(u2, N) % N
    if s > N // 2:  # negating both scalars keeps x(R) and flips s
        s = N - s
        u1 = (N - u1) % N
    m = u1 * s % N
    b = lambda v: np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8)
    pub33 = np.concatenate([np.array([2 + (Q[1] & 1)], np.uint8), b(Q[0])])
    return b(m).copy(), pub33, np.concatenate([b(Q[0]), b(Q[1])]), np.concatenate([b(r), b(s)])


def special_scalars():
    a1 = 0x3086D221A7D46BCDE86C90E49284EB15
    b1 = 0xE4437ED6010E88286F547FA90ABFE4C3
    a2 = 0x114CA50F7A8E2F3F657C1108D9D44CFD8
    vals = [0, 1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 33, N - 1, N - 2, N - 3, (N - 1) // 2, (N + 1) // 2,
            LAMBDA, LAMBDA + 1, LAMBDA - 1, N - LAMBDA, 2 * LAMBDA % N, a1, b1, a2, N - a1, N - b1,
            (a1 + b1 * LAMBDA) % N, 2**128, 2**128 - 1, 2**128 + 1, 2**127, 2**129, 2**255, 2**255 - 1,
            0x8000, 0x8001, 0x7FFF, 0xFFFF, 0x10000, 0x80000000, 0xFFFF << 240, 0x8000 << 240, (0x8000 << 224) - 1,
            int("8000" * 16, 16) % N, int("7FFF" * 16, 16), int("8001" * 16, 16) % N, int("F" * 64, 16) % N,
            int("1" * 64, 16), int("8" * 64, 16) % N]
    return vals


def cases(limit=None):
    """(msg32[n,32], pub33[n,33], pubxy[n,64], sig[n,64]) arrays of crafted signatures."""
    sc = special_scalars()
    out = []
    ds = [1, 2, 3, 5, N - 1, N - 2, (N + 1) // 2, LAMBDA, 0xDEADBEEFCAFEBABE0123456789ABCDEF]
    for di, d in enumerate(ds):
        for i, u2 in enumerate(sc):
            for j, u1 in enumerate(sc):
                if (i + 2 * j + di) % (3 if di < 4 else 11):  # thin the cross product deterministically
                    continue
                c = craft(d, u1, u2)
                if c is not None:
                    out.append(c)
    # collisions inside the ladder: u2 small multiples with Q = G so that u1*G and u2*Q cancel or coincide
    for u in range(1, 40):
        for d in (1, 2, N - 1):
            for u1 in (u, N - u, (u * d) % N, (N - u * d) % N, (2 * u * d) % N):
                c = craft(d, u1, u)
                if c is not None:
                    out.append(c)
    if limit:
        out = out[:limit]
    cols = list(zip(*out))
    return tuple(np.stack(c) for c in cols)
