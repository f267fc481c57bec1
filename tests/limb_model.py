"""Python-integer model of the engine's device limb algorithms (fe.cuh, sc.cuh), step by step.

Used only by tests.  Each function mirrors the DEVICE code path (the inline-PTX one, `#if SV_DEVICE_CODE`), returns
the exact limbs the device must produce and says which rare branch the operands take, so that the GPU self test
(tests/test_gpu_selftest.py) can (a) compare raw outputs bit for bit, not only modulo p, and (b) report how often each
rare branch was exercised.  Constructors build operands that reach every branch on purpose.
"""
import random

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
C = 2**32 + 977  # 2^256 mod p
NC = 2**256 - N  # 2^256 mod n (129 bits)
M256 = 2**256 - 1
M64 = 2**64 - 1


# ---------------------------------------------------------------------------------------------------------------
# field: fe_add / fe_sub / fe_reduce512 exactly as fe.cuh's device branches compute them
# ---------------------------------------------------------------------------------------------------------------
def fe_add(a, b):
    """-> (raw result, flags) ; flags: c (carry of the 256-bit add), k (ripple past limb 1), c2 (second wrap)"""
    s = a + b
    c = s >> 256
    r = s & M256
    lo = (r & M64) + c * C  # mad.lo.cc / addc.cc on limbs 0,1 ; k = carry out of limb 1
    k = lo >> 64
    r = (r & ~M64 & M256) | (lo & M64)
    c2 = 0
    if k:
        hi = (r >> 64) + 1  # ripple through limbs 2..7
        c2 = hi >> 192
        r = ((hi & (2**192 - 1)) << 64) | (r & M64)
        lo = (r & M64) + c2 * C  # limbs 0,1 only; cannot carry further (value < 2^34 when c2)
        assert lo >> 64 == 0 or not c2
        r = (r & ~M64 & M256) | (lo & M64)
    return r, dict(c=c, k=k, c2=c2)


def fe_sub(a, b):
    d = a - b
    bw = 1 if d < 0 else 0
    r = d & M256
    lo = (r & M64) - bw * C
    k = 1 if lo < 0 else 0
    r = (r & ~M64 & M256) | (lo & M64)
    b2 = 0
    if k:
        hi = (r >> 64) - 1
        b2 = 1 if hi < 0 else 0
        r = ((hi & (2**192 - 1)) << 64) | (r & M64)
        lo = (r & M64) - b2 * C
        r = (r & ~M64 & M256) | (lo & M64)
    return r, dict(bw=bw, k=k, b2=b2)


def fe_reduce512(t):
    lo, hi = t & M256, t >> 256
    S = lo + hi * C  # 10 limbs, exact
    T = S >> 256
    s = S & M256
    f = T * C  # three limbs
    r = s + f
    c = r >> 256
    r &= M256
    r += c * C  # third fold: limbs 0..2, cannot wrap
    assert r <= M256
    return r, dict(T=T, c=c)


def fe_mul(a, b):
    return fe_reduce512(a * b)


def fe_sqr(a):
    return fe_reduce512(a * a)


def fe_mul8(a):
    top = a >> 253
    x = (a << 3) & M256
    return fe_add(x, top * C)


def fe_mul3(a):
    t, f1 = fe_add(a, a)
    r, f2 = fe_add(t, a)
    return r, dict(k=f1["k"] | f2["k"], c2=f1["c2"] | f2["c2"])


def fe_normalize(a):
    return a - P if a >= P else a


# ---------------------------------------------------------------------------------------------------------------
# constructors of rare-branch operands
# ---------------------------------------------------------------------------------------------------------------
def add_k_operands(rnd, second_wrap=False):
    """(a, b) with a + b >= 2^256 whose low 64 bits of the wrapped sum are >= 2^64 - C: the carry fold ripples past
    limb 1.  second_wrap: limbs 2..7 of the wrapped sum are all ones too, so the ripple wraps a second time."""
    low = rnd.randrange(2**64 - C, 2**64)
    mid = (2**192 - 1) if second_wrap else rnd.getrandbits(192)
    if not second_wrap and mid == 2**192 - 1:
        mid -= 1
    r = (mid << 64) | low  # wrapped sum
    total = 2**256 + r
    a = rnd.randrange(total - M256, M256 + 1)  # both a and b must fit 256 bits
    return a, total - a


def sub_k_operands(rnd, second_wrap=False):
    """(a, b) with a < b and low 64 bits of a - b + 2^256 below C (borrow past limb 1); second_wrap: limbs 2..7 zero."""
    low = rnd.randrange(0, C)
    mid = 0 if second_wrap else rnd.randrange(1, 2**192)
    r = (mid << 64) | low  # a - b + 2^256
    # a - b = r - 2^256 ; choose b in (2^256 - r .. M256], a = b + r - 2^256 >= 0
    b = rnd.randrange(2**256 - r, M256 + 1)
    return b + r - 2**256, b


def reduce_third_fold_t(rnd):
    """512-bit t whose second fold s[0..7] + T*C carries out of 2^256 (fe_reduce512's `c`)."""
    while True:
        hi = rnd.getrandbits(256) | (1 << 255)
        T_guess = (hi * C) >> 256
        if T_guess == 0:
            continue
        delta = rnd.randrange(1, T_guess * C)
        lo = (2**256 - delta - hi * C) % 2**256
        t = (hi << 256) | lo
        if fe_reduce512(t)[1]["c"]:
            return t


def mul_third_fold_operands(rnd, square=False):
    """(a, b) (or a alone) whose product takes the third fold: a*b mod p = w with C <= w < (T+1)*C."""
    while True:
        w = rnd.randrange(C, 2**62)
        if square:
            if pow(w, (P - 1) // 2, P) != 1:
                continue
            a = pow(w, (P + 1) // 4, P)
            if rnd.getrandbits(1):
                a = P - a
            if fe_sqr(a)[1]["c"]:
                return a, a
        else:
            a = rnd.randrange(2**255, P)
            b = w * pow(a, P - 2, P) % P
            if fe_mul(a, b)[1]["c"]:
                return a, b


# ---------------------------------------------------------------------------------------------------------------
# scalar: sc_reduce512 as sc.cuh computes it (four folds + one conditional subtraction)
# ---------------------------------------------------------------------------------------------------------------
def sc_reduce512(t):
    m = (t & M256) + (t >> 256) * NC  # fold 1: 13..14 limbs
    q = (m & M256) + (m >> 256) * NC  # fold 2
    q8 = q >> 256
    assert q8 < 2**6
    s = (q & M256) + q8 * NC  # fold 3
    c = s >> 256
    s &= M256
    s += c * NC  # fold 4 (value tiny when c)
    assert s <= M256
    sub = s >= N
    if sub:
        s -= N
    assert s < N, "one conditional subtraction must suffice"
    return s, dict(q8=q8, c=c, sub=sub)


def sc_fold4_t(rnd):
    """512-bit t taking the carry of the third fold (c = 1), built backwards through the folds: the second fold must land
    in [2^257 - NC, 2^257)."""
    while True:
        q = 2**257 - rnd.randrange(1, NC + 1)
        mh = rnd.randrange(2**256 // NC + 1, NC)
        ml = q - mh * NC
        if not 0 <= ml < 2**256:
            continue
        m = (mh << 256) | ml
        th = m // NC
        if th > M256:
            continue
        t = (th << 256) | (m - th * NC)
        if sc_reduce512(t)[1]["c"]:
            return t


def sc_mul_fold4_operands(rnd):
    """(a, b), both < n, whose product takes that carry: a*b mod n in [NC, 2NC) reached as the representative + 2n."""
    while True:
        r = rnd.randrange(NC, 2 * NC)
        a = rnd.randrange(1, N)
        b = r * pow(a, N - 2, N) % N
        if sc_reduce512(a * b)[1]["c"]:
            return a, b


def split_lambda(k):
    """secp256k1_scalar_split_lambda (scalar_impl.h:138-176) with Python integers."""
    g1 = 0x3086D221A7D46BCDE86C90E49284EB153DAA8A1471E8CA7FE893209A45DBB031
    g2 = 0xE4437ED6010E88286F547FA90ABFE4C4221208AC9DF506C61571B4AE8AC47F71
    mb1 = 0xE4437ED6010E88286F547FA90ABFE4C3
    mb2 = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFE8A280AC50774346DD765CDA83DB1562C
    lam = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
    c1 = (k * g1 + (1 << 383)) >> 384
    c2 = (k * g2 + (1 << 383)) >> 384
    r2 = (c1 * mb1 + c2 * mb2) % N
    r1 = (k - r2 * lam) % N
    return r1, r2


EDGE_FE = [0, 1, 2, 3, 7, 8, 977, C - 1, C, C + 1, 2 * C, 2**32 - 1, 2**32, 2**64 - 1, 2**64, 2**64 - C, 2**64 - C - 1,
           2**128 - 1, 2**128, 2**192, 2**224, 2**255, 2**255 - 1, 2**255 + 1, (P + 1) // 2, (P - 1) // 2, P - C, P - 2,
           P - 1, P, P + 1, P + 2, P + C - 2, 2**256 - C, 2**256 - C - 1, 2**256 - 2**64, 2**256 - 2**32 - 978,
           2**256 - 2, 2**256 - 1, (2**256 - 1) ^ (2**128 - 1), (2**256 - 1) ^ (2**64 - 1), 0x5555555555555555555555555555555555555555555555555555555555555555,
           0xAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA,
           0xFFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000,
           0x00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF,
           0x8000000080000000800000008000000080000000800000008000000080000000,
           0x7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF7FFFFFFF]

EDGE_SC = [0, 1, 2, 3, N - 1, N - 2, N - 3, (N - 1) // 2, (N + 1) // 2, (N - 1) // 2 - 1, 2**128, 2**128 - 1, 2**129, 2**255,
           2**255 - 1, NC, NC - 1, NC + 1, N - NC, 2**64 - 1, 2**192,
           0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72,
           N - 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72,
           0x3086D221A7D46BCDE86C90E49284EB15, 0xE4437ED6010E88286F547FA90ABFE4C3]


def limb_patterns(rnd, count):
    """256-bit values whose 32-bit limbs are drawn from carry-stressing constants"""
    pool = [0, 1, 2, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF, 0x00010000, 0x0000FFFF, 0xFFFF0000]
    out = []
    for _ in range(count):
        v = 0
        for i in range(8):
            limb = rnd.choice(pool) if rnd.random() < 0.8 else rnd.getrandbits(32)
            v |= limb << (32 * i)
        out.append(v)
    return out
