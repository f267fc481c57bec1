"""Self-test cases: the arithmetic of the engine, primitive by primitive, against Python integers.

Run on the GPU through the C ABI by tests/test_gpu_selftest.py (the inline-PTX device forms) and, at reduced sizes, on the
host build of the same headers by tests/test_selftest_emul.py.

Model: libsecp256k1's own self tests (tests.c:3023-3176 field, :2354 scalar).  The device code path of u256_add/sub,
sv_mul8_dev / sv_sqr8_dev (generated PTX), fe_reduce512, fe_add / fe_sub (rare carry folds) and the scalar folds is a
different implementation from the uint64 fallback tests/host_emul exercises, so it is pinned here, on the hardware,
through the C ABI (sv_selftest_host):
  * an edge table (all pairs) and carry-stressing limb patterns,
  * operands CONSTRUCTED to take every rare branch (tests/limb_model.py predicts the branch and the exact limbs),
  * >= 10^7 random pairs for fe_mul, 2x10^6 for the other binary ops.
The hit count of every rare branch is written to gpurun_out/selftest_coverage.json (committed under profiles/).
"""
import json
import os
import random

import numpy as np

from tests import limb_model as M

P, N = M.P, M.N
OPS = dict(FE_MUL=0, FE_SQR=1, FE_ADD=2, FE_SUB=3, FE_NEG=4, FE_NORMALIZE=5, FE_INV=6, FE_SQRT=7, FE_MUL3=8, FE_MUL8=9,
           FE_MUL_SMALL=10, FE_DBL=11, FE_B32=12, U256_MUL_WIDE=13, U256_SQR_WIDE=14, FE_REDUCE512=15, U256_ADD=16,
           U256_SUB=17, SC_MUL=20, SC_SQR=21, SC_ADD=22, SC_NEGATE=23, SC_INVERSE=24, SC_REDUCE512=25, SC_SPLIT_LAMBDA=26,
           SC_SET_B32=27, ECMULT_GEN=28, PREPARE_U2=29, PREPARE_U1=30, SC_INVERSE_VAR=31, FE_INV_VAR=32)
COVERAGE = {}


def limbs(vals):
    """list of ints (< 2^256) -> (n, 8) uint32 little-endian limbs"""
    buf = b"".join(v.to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint32).reshape(-1, 8).copy()


def ints(arr, words=8):
    """(n, >=words) uint32 -> list of ints from the first `words` limbs"""
    a = np.ascontiguousarray(arr[:, :words])
    mv = memoryview(a.tobytes())
    w = 4 * words
    return [int.from_bytes(mv[i * w:(i + 1) * w], "little") for i in range(a.shape[0])]


def hit(name, flags):
    for k, v in flags.items():
        if v:
            COVERAGE[f"{name}.{k}"] = COVERAGE.get(f"{name}.{k}", 0) + 1


def write_coverage(name):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(dict(sorted(COVERAGE.items())), f, indent=1)


def run2(engine, op, A, B):
    return engine.selftest(OPS[op], limbs(A), limbs(B))


# ---------------------------------------------------------------------------------------------------------------
def case_u256_products_and_add_sub_exact(engine):
    rnd = random.Random(11)
    vals = M.EDGE_FE + M.limb_patterns(rnd, 600) + [rnd.getrandbits(256) for _ in range(200)]
    A, B = [], []
    for a in vals:
        for b in rnd.sample(vals, 24) + M.EDGE_FE[-12:]:
            A.append(a)
            B.append(b)
    out = run2(engine, "U256_MUL_WIDE", A, B)
    got = ints(out, 16)
    assert all(g == a * b for g, a, b in zip(got, A, B)), "sv_mul8_dev: 512-bit product differs"
    got = ints(engine.selftest(OPS["U256_SQR_WIDE"], limbs(vals)), 16)
    assert all(g == a * a for g, a in zip(got, vals)), "sv_sqr8_dev: 512-bit square differs"
    out = run2(engine, "U256_ADD", A, B)
    assert all(g == (a + b) % 2**256 and int(c) == (a + b) >> 256 for g, c, a, b in zip(ints(out), out[:, 8], A, B))
    out = run2(engine, "U256_SUB", A, B)
    assert all(g == (a - b) % 2**256 and int(c) == (1 if a < b else 0) for g, c, a, b in zip(ints(out), out[:, 8], A, B))
    COVERAGE["u256.pairs"] = len(A)


def case_field_edge_all_pairs_raw_limbs(engine):
    """every pair of the edge table through fe_mul/add/sub: the RAW device limbs equal the step-by-step model (so every
    branch taken is the predicted one), and the value is right modulo p"""
    E = M.EDGE_FE
    A = [a for a in E for _ in E]
    B = [b for _ in E for b in E]
    for op, model, exact in (("FE_MUL", M.fe_mul, lambda a, b: a * b), ("FE_ADD", M.fe_add, lambda a, b: a + b),
                             ("FE_SUB", M.fe_sub, lambda a, b: a - b)):
        got = ints(run2(engine, op, A, B))
        for g, a, b in zip(got, A, B):
            r, f = model(a, b)
            assert g == r, (op, hex(a), hex(b), hex(g), hex(r), f)
            assert g % P == exact(a, b) % P
            hit(op.lower() + ".edge", f)


def case_field_unary_ops_edge(engine):
    rnd = random.Random(12)
    vals = M.EDGE_FE + M.limb_patterns(rnd, 300) + [rnd.getrandbits(256) for _ in range(300)]
    Z = [0] * len(vals)
    got = ints(run2(engine, "FE_SQR", vals, Z))
    for g, a in zip(got, vals):
        r, f = M.fe_sqr(a)
        assert g == r and g % P == a * a % P
        hit("fe_sqr.edge", f)
    for g, a in zip(ints(run2(engine, "FE_NEG", vals, Z)), vals):
        assert g == M.fe_sub(0, a)[0] and g % P == (-a) % P
    for g, a in zip(ints(run2(engine, "FE_DBL", vals, Z)), vals):
        assert g == M.fe_add(a, a)[0] and g % P == 2 * a % P
    for g, a in zip(ints(run2(engine, "FE_MUL3", vals, Z)), vals):
        assert g == M.fe_mul3(a)[0] and g % P == 3 * a % P
    for g, a in zip(ints(run2(engine, "FE_MUL8", vals, Z)), vals):
        assert g == M.fe_mul8(a)[0] and g % P == 8 * a % P
    ks = [rnd.choice([0, 1, 2, 3, 7, 8, 21, 977, 65535]) for _ in vals]
    for g, a, k in zip(ints(run2(engine, "FE_MUL_SMALL", vals, ks)), vals, ks):
        assert g % P == k * a % P
    B = [rnd.choice([a, a % P, (a + P) % 2**256 if a < M.C else a, (a + 1) % 2**256]) for a in vals]
    out = run2(engine, "FE_NORMALIZE", vals, B)
    for g, fl, a, b in zip(ints(out), out, vals, B):
        assert g == a % P and int(fl[8]) == (a % P == 0) and int(fl[9]) == (a >= P) and int(fl[10]) == (a % P == b % P)
    # big-endian byte import/export: limbs hold the 32 bytes in memory order
    be = [int.from_bytes(v.to_bytes(32, "big"), "little") for v in vals]
    out = run2(engine, "FE_B32", be, Z)
    for g, fl, v in zip(ints(out), out, vals):
        assert int(fl[8]) == (v < P)
        assert int.from_bytes(g.to_bytes(32, "little"), "big") == v % P
    sub = vals[:160]
    for g, a in zip(ints(run2(engine, "FE_INV", sub, [0] * len(sub))), sub):
        assert g % P == pow(a, P - 2, P)
    out = run2(engine, "FE_SQRT", sub, [0] * len(sub))
    for g, fl, a in zip(ints(out), out, sub):
        qr = pow(a % P, (P - 1) // 2, P) in (0, 1)
        assert int(fl[8]) == qr
        if qr:
            assert g * g % P == a % P
    # squares always have roots
    sq = [v * v % P for v in vals[:100]]
    out = run2(engine, "FE_SQRT", sq, [0] * len(sq))
    assert all(int(fl[8]) == 1 and g * g % P == a for g, fl, a in zip(ints(out), out, sq))


def case_field_rare_branches_constructed(engine):
    """operands built to take each rare-carry branch of the device code: fe_add k-path and its second wrap, fe_sub k-path
    and its second wrap, the third fold of fe_reduce512 reached directly, through fe_mul and through fe_sqr"""
    rnd = random.Random(13)
    n = 4000
    for name, gen, op, model in (
            ("fe_add.k", lambda: M.add_k_operands(rnd, False), "FE_ADD", M.fe_add),
            ("fe_add.k+wrap", lambda: M.add_k_operands(rnd, True), "FE_ADD", M.fe_add),
            ("fe_sub.k", lambda: M.sub_k_operands(rnd, False), "FE_SUB", M.fe_sub),
            ("fe_sub.k+wrap", lambda: M.sub_k_operands(rnd, True), "FE_SUB", M.fe_sub)):
        pairs = [gen() for _ in range(n)]
        A, B = [p[0] for p in pairs], [p[1] for p in pairs]
        exact = (lambda a, b: a + b) if op == "FE_ADD" else (lambda a, b: a - b)
        for g, a, b in zip(ints(run2(engine, op, A, B)), A, B):
            r, f = model(a, b)
            assert f["k"] == 1 and (("wrap" not in name) or f.get("c2", f.get("b2")) == 1)
            assert g == r and g % P == exact(a, b) % P, (name, hex(a), hex(b))
            hit(name.split(".")[0] + ".constructed", f)
    ts = [M.reduce_third_fold_t(rnd) for _ in range(n)]
    got = ints(run2(engine, "FE_REDUCE512", [t & M.M256 for t in ts], [t >> 256 for t in ts]))
    for g, t in zip(got, ts):
        r, f = M.fe_reduce512(t)
        assert f["c"] == 1 and g == r and g % P == t % P
        hit("fe_reduce512.constructed", f)
    pairs = [M.mul_third_fold_operands(rnd) for _ in range(1500)]
    A, B = [p[0] for p in pairs], [p[1] for p in pairs]
    for g, a, b in zip(ints(run2(engine, "FE_MUL", A, B)), A, B):
        r, f = M.fe_mul(a, b)
        assert f["c"] == 1 and g == r and g % P == a * b % P
        hit("fe_mul.constructed", f)
    A = [M.mul_third_fold_operands(rnd, square=True)[0] for _ in range(1500)]
    for g, a in zip(ints(run2(engine, "FE_SQR", A, [0] * len(A))), A):
        r, f = M.fe_sqr(a)
        assert f["c"] == 1 and g == r and g % P == a * a % P
        hit("fe_sqr.constructed", f)
    # generic reduce512 inputs: extremes of the 512-bit range
    ts = [0, 1, 2**512 - 1, 2**256, 2**256 - 1, P * P, (P - 1) ** 2, (2**256 - 1) ** 2, 2**512 - 2**256, P << 256, (P << 256) | P] + \
         [rnd.getrandbits(512) for _ in range(3000)]
    got = ints(run2(engine, "FE_REDUCE512", [t & M.M256 for t in ts], [t >> 256 for t in ts]))
    for g, t in zip(got, ts):
        r, f = M.fe_reduce512(t)
        assert g == r and g % P == t % P
        hit("fe_reduce512.generic", f)


def _random_pairs_check(engine, op, total, chunk, check, seed, weak_frac=0.02):
    rng = np.random.default_rng(seed)
    done = 0
    while done < total:
        m = min(chunk, total - done)
        a = rng.integers(0, 2**32, size=(m, 8), dtype=np.uint32)
        b = rng.integers(0, 2**32, size=(m, 8), dtype=np.uint32)
        k = int(m * weak_frac)  # a slice of operands in the non-canonical top range [p, 2^256) and near 0 / 2^256
        a[:k, 2:] = 0xFFFFFFFF
        b[k:2 * k, 1:] = 0xFFFFFFFF
        a[2 * k:3 * k, 1:] = 0
        out = engine.selftest(OPS[op], a, b)
        A, B, R = ints(a), ints(b), ints(out)
        bad = [i for i in range(m) if not check(R[i], A[i], B[i])]
        assert not bad, (op, done + bad[0], hex(A[bad[0]]), hex(B[bad[0]]), hex(R[bad[0]]))
        done += m
    COVERAGE[f"{op.lower()}.random_pairs"] = COVERAGE.get(f"{op.lower()}.random_pairs", 0) + total


def case_field_mul_ten_million_random_pairs(engine, total=10_000_000):
    _random_pairs_check(engine, "FE_MUL", total, min(total, 1_000_000), lambda r, a, b: r % P == a * b % P, seed=101)


def case_field_other_ops_random_pairs(engine, m=1_000_000):
    _random_pairs_check(engine, "FE_SQR", 2 * m, m, lambda r, a, b: r % P == a * a % P, seed=102)
    _random_pairs_check(engine, "FE_ADD", 2 * m, m, lambda r, a, b: r % P == (a + b) % P, seed=103)
    _random_pairs_check(engine, "FE_SUB", 2 * m, m, lambda r, a, b: r % P == (a - b) % P, seed=104)
    _random_pairs_check(engine, "U256_MUL_WIDE", m, m, lambda r, a, b: True, seed=105)  # shape only; exact below
    rng = np.random.default_rng(106)
    a = rng.integers(0, 2**32, size=(m, 8), dtype=np.uint32)
    b = rng.integers(0, 2**32, size=(m, 8), dtype=np.uint32)
    out = engine.selftest(OPS["U256_MUL_WIDE"], a, b)
    assert all(r == x * y for r, x, y in zip(ints(out, 16), ints(a), ints(b)))
    out = engine.selftest(OPS["U256_SQR_WIDE"], a)
    assert all(r == x * x for r, x in zip(ints(out, 16), ints(a)))
    COVERAGE["u256_mul_wide.random_pairs_exact"] = m
    COVERAGE["u256_sqr_wide.random_exact"] = m


# ---------------------------------------------------------------------------------------------------------------
def case_scalar_ops_edge_and_rare_folds(engine):
    rnd = random.Random(21)
    E = M.EDGE_SC + [rnd.randrange(N) for _ in range(40)]
    A = [a for a in E for _ in E]
    B = [b for _ in E for b in E]
    for g, a, b in zip(ints(run2(engine, "SC_MUL", A, B)), A, B):
        assert g == a * b % N
        hit("sc_mul.edge", {k: v for k, v in M.sc_reduce512(a * b)[1].items() if k != "sub"})
    for g, a, b in zip(ints(run2(engine, "SC_ADD", A, B)), A, B):
        assert g == (a + b) % N
    Z = [0] * len(E)
    for g, a in zip(ints(run2(engine, "SC_SQR", E, Z)), E):
        assert g == a * a % N
    out = run2(engine, "SC_NEGATE", E, Z)
    for g, fl, a in zip(ints(out), out, E):
        assert g == (-a) % N and int(fl[8]) == (a > (N - 1) // 2) and int(fl[9]) == (a == 0) and int(fl[10]) == 0
    for g, a in zip(ints(run2(engine, "SC_INVERSE", E, Z)), E):
        assert g == pow(a, N - 2, N)
    # the variable-time inverses of the small-batch path (binary extended Euclid): edge values, powers of two and their
    # neighbours (long shift runs), random values
    inv_in = E + [2**k for k in range(1, 256, 7)] + [2**k - 1 for k in range(2, 256, 11)] + [N - 2**k for k in range(1, 250, 13)] + \
        [rnd.randrange(1, N) for _ in range(1500)]
    for g, a in zip(ints(run2(engine, "SC_INVERSE_VAR", inv_in, [0] * len(inv_in))), inv_in):
        assert g == pow(a, N - 2, N), hex(a)
    fin = [v for v in M.EDGE_FE] + [2**k for k in range(1, 256, 9)] + [P - 2**k for k in range(1, 250, 17)] + [rnd.getrandbits(256) for _ in range(1500)]
    for g, a in zip(ints(run2(engine, "FE_INV_VAR", fin, [0] * len(fin))), fin):
        assert g == pow(a % P, P - 2, P), hex(a)
    # the carry of the third fold, reached directly and through products of reduced scalars
    ts = [M.sc_fold4_t(rnd) for _ in range(3000)] + [2**512 - 1, 0, 2**256, N * N, (N - 1) ** 2, 2**512 - 2**256, N << 256] + \
         [rnd.getrandbits(512) for _ in range(3000)]
    got = ints(run2(engine, "SC_REDUCE512", [t & M.M256 for t in ts], [t >> 256 for t in ts]))
    for i, (g, t) in enumerate(zip(got, ts)):
        r, f = M.sc_reduce512(t)
        assert g == r == t % N
        assert i >= 3000 or f["c"] == 1
        hit("sc_reduce512", f)
    pairs = [M.sc_mul_fold4_operands(rnd) for _ in range(1500)]
    A, B = [p[0] for p in pairs], [p[1] for p in pairs]
    for g, a, b in zip(ints(run2(engine, "SC_MUL", A, B)), A, B):
        assert g == a * b % N and M.sc_reduce512(a * b)[1]["c"] == 1
        hit("sc_mul.constructed", {"c": 1})
    # set_b32 with overflow: values >= n must reduce and report it (scalar_4x64_impl.h:158-170)
    vals = [0, 1, N - 1, N, N + 1, 2**256 - 1, 2**256 - 2, N + M.NC - 1] + [rnd.getrandbits(256) for _ in range(500)] + \
           [N + rnd.randrange(M.NC) for _ in range(500)]
    be = [int.from_bytes(v.to_bytes(32, "big"), "little") for v in vals]
    out = run2(engine, "SC_SET_B32", be, [0] * len(be))
    for g, fl, v in zip(ints(out), out, vals):
        assert g == v % N and int(fl[8]) == (v >= N)
    out = run2(engine, "SC_NEGATE", [v for v in vals if v < 2**256], [0] * len(vals))
    for fl, v in zip(out, vals):
        assert int(fl[10]) == (v >= N)


def case_scalar_mul_two_million_random_pairs(engine, m=1_000_000):
    rng = np.random.default_rng(201)
    for _ in range(2):
        a = rng.integers(0, 2**32, size=(m, 8), dtype=np.uint32)
        b = rng.integers(0, 2**32, size=(m, 8), dtype=np.uint32)
        a[:, 7] >>= 1  # < 2^255 < n: sc_mul's contract is reduced operands
        b[:, 7] >>= 1
        out = engine.selftest(OPS["SC_MUL"], a, b)
        assert all(r == x * y % N for r, x, y in zip(ints(out), ints(a), ints(b)))
    COVERAGE["sc_mul.random_pairs"] = 2 * m


def case_glv_split_and_recoding_on_device(engine):
    """scalar_split_lambda (scalar_impl.h:138-176; bounds tests.c:4560,5805) and the engine's own recodings, device side"""
    from tests import adversarial
    rnd = random.Random(22)
    ks = [v % N for v in adversarial.special_scalars()] + M.EDGE_SC + [rnd.randrange(N) for _ in range(20000)]
    out = engine.selftest(OPS["SC_SPLIT_LAMBDA"], limbs(ks))
    lam = adversarial.LAMBDA
    for r1, r2, k in zip(ints(out), ints(out[:, 8:]), ks):
        w1, w2 = M.split_lambda(k)
        assert (r1, r2) == (w1, w2)
        assert (r1 + r2 * lam - k) % N == 0
        assert min(r1, N - r1) < 2**128 and min(r2, N - r2) < 2**128
    out = engine.selftest(OPS["PREPARE_U2"], limbs([0] * len(ks)), limbs(ks))
    for row, k in zip(out, ks):
        a = sum(int(row[i]) << (32 * i) for i in range(5))
        b = sum(int(row[5 + i]) << (32 * i) for i in range(5))
        sa, sb = a >> 159, b >> 159
        a &= (1 << 159) - 1
        b &= (1 << 159) - 1
        assert a & 1 and b & 1 and a < 2**131 and b < 2**131
        assert ((-a if sa else a) + (-b if sb else b) * lam - k) % N == 0
    out = engine.selftest(OPS["PREPARE_U1"], limbs(ks)).astype(np.int32)
    for row, k in zip(out, ks):
        assert sum(int(row[i]) << (16 * i) for i in range(16)) == k
        assert all(-32768 <= int(row[i]) <= 32768 for i in range(15)) and 0 <= int(row[15]) <= 65536


def case_ecmult_kat_through_device_comb_table(engine):
    """tests.c:5657-5726 (test_ecmult_constants_sha): SHA-256 over the uncompressed serialisation of x*G for
    x in {0, 1, -1, SHA256(LE32(prefix) || LE16(i))}, infinity hashed as one zero byte — computed here by the ENGINE's
    fixed-base comb (the 34 MiB device table built by k_gtable_fill), not by the port."""
    import hashlib
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ecmult_kat.json")
    for kat in json.load(open(gold)):
        scalars = [0, 1, N - 1]
        for i in range(kat["iters"]):
            inp = kat["prefix"].to_bytes(4, "little") + i.to_bytes(2, "little")
            scalars.append(int.from_bytes(hashlib.sha256(inp).digest(), "big") % N)
        out = engine.selftest(OPS["ECMULT_GEN"], limbs(scalars))
        acc = hashlib.sha256()
        for x, px, py in zip(scalars, ints(out), ints(out[:, 8:])):
            acc.update(b"\x00" if x == 0 else b"\x04" + px.to_bytes(32, "big") + py.to_bytes(32, "big"))
        assert acc.hexdigest() == kat["sha256"], kat
    COVERAGE["ecmult_kat.digests"] = 2
