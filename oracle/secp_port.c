/*
 * oracle/secp_port.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement ("port") of the reference's verification path, written from the
 * algorithm descriptions in the reference sources (cited per function, paths relative to
 * /root/reference/external/libwally-core/src/secp256k1/src unless noted).  It exists so that the
 * CUDA engine can be checked on a machine where /root/reference is absent, and so that the
 * accept/reject rules are stated once, compactly, in reviewable C.
 *
 * PARITY PINNED: tests/test_oracle.py checks this file against (1) the unmodified reference
 * compiled by oracle/Makefile (oracle/_ref), on random + corrupted + adversarial inputs, and
 * (2) the golden vectors extracted from the reference's own tests (tests/golden/: Wycheproof
 * ECDSA 463 vectors, BIP-340 vectors 0-14, the gossip_store fixture, ...).
 *
 * Simplifications that cannot change a verdict: 4x64-bit fully-reduced field/scalar limbs
 * instead of 5x52 lazy limbs (field_5x52.h) ; Fermat inversions instead of safegcd (modinv64) ;
 * Strauss double-scalar multiplication with width-5 wNAF for BOTH points and no endomorphism
 * split (ecmult_impl.h:234-341 uses w=5 + GLV for A and 2x8192-entry static tables for G).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint8_t u8;

/* ------------------------------------------------------------------ 256-bit helpers */
typedef struct { u64 d[4]; } num; /* little-endian limbs */

static int num_cmp(const num *a, const num *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->d[i] < b->d[i]) return -1;
        if (a->d[i] > b->d[i]) return 1;
    }
    return 0;
}
static int num_is_zero(const num *a) { return (a->d[0] | a->d[1] | a->d[2] | a->d[3]) == 0; }
static u64 num_add(num *r, const num *a, const num *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->d[i] + b->d[i]; r->d[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static u64 num_sub(num *r, const num *a, const num *b) {
    u64 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a->d[i] - b->d[i] - bw;
        r->d[i] = (u64)t;
        bw = (u64)(t >> 64) & 1;
    }
    return bw;
}
static void num_from_be(num *r, const u8 *b) {
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (3 - i) + j];
        r->d[i] = v;
    }
}
static void num_to_be(u8 *b, const num *a) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (u8)(a->d[i] >> (56 - 8 * j));
}
static void mul_wide(u64 t[8], const num *a, const num *b) {
    memset(t, 0, 64);
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->d[j] * b->d[i] + t[i + j];
            t[i + j] = (u64)c;
            c >>= 64;
        }
        t[i + 4] = (u64)c;
    }
}

/* ------------------------------------------------------------------ field F_p (field.h, field_impl.h) */
static const num FE_P = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
#define FE_C 0x1000003D1ULL /* 2^256 mod p (field_5x52_impl.h:482) */
typedef num fe; /* always canonical: < p */

static void fe_fix(fe *r) { /* single conditional subtraction */
    if (num_cmp(r, &FE_P) >= 0) num_sub(r, r, &FE_P);
}
/* secp256k1_fe_set_b32_limit (field_5x52_impl.h:272): fails for values >= p */
static int fe_set_b32_limit(fe *r, const u8 *b) { num_from_be(r, b); return num_cmp(r, &FE_P) < 0; }
static void fe_get_b32(u8 *b, const fe *a) { num_to_be(b, a); }
static void fe_add(fe *r, const fe *a, const fe *b) {
    u64 c = num_add(r, a, b);
    if (c) { num k = {{FE_C, 0, 0, 0}}; num_add(r, r, &k); } /* wrapped value is < p, + C cannot wrap */
    fe_fix(r);
}
static void fe_neg(fe *r, const fe *a) { /* secp256k1_fe_negate */
    if (num_is_zero(a)) { *r = *a; return; }
    num_sub(r, &FE_P, a);
}
static void fe_reduce(fe *r, const u64 t[8]) {
    /* lo + hi * C, twice */
    u64 s[5];
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)t[i] + (u128)t[4 + i] * FE_C; s[i] = (u64)c; c >>= 64; }
    s[4] = (u64)c; /* < 2^34 */
    c = (u128)s[0] + (u128)s[4] * FE_C;
    r->d[0] = (u64)c; c >>= 64;
    for (int i = 1; i < 4; i++) { c += s[i]; r->d[i] = (u64)c; c >>= 64; }
    if (c) { num k = {{FE_C, 0, 0, 0}}; num_add(r, r, &k); }
    fe_fix(r);
}
static void fe_mul(fe *r, const fe *a, const fe *b) { u64 t[8]; mul_wide(t, a, b); fe_reduce(r, t); }
static void fe_sqr(fe *r, const fe *a) { fe_mul(r, a, a); }
static void fe_set_int(fe *r, u64 v) { r->d[0] = v; r->d[1] = r->d[2] = r->d[3] = 0; }
static int fe_equal(const fe *a, const fe *b) { return num_cmp(a, b) == 0; }
static int fe_is_odd(const fe *a) { return (int)(a->d[0] & 1); }
static void fe_pow(fe *r, const fe *a, const num *e) { /* plain square-and-multiply */
    fe acc; fe_set_int(&acc, 1);
    for (int i = 255; i >= 0; i--) {
        fe_sqr(&acc, &acc);
        if ((e->d[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}
/* secp256k1_fe_sqrt (field_impl.h:33-138): a^((p+1)/4), then check by squaring */
static int fe_sqrt(fe *r, const fe *a) {
    static const num E = {{0xFFFFFFFFBFFFFF0CULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFFFFULL}};
    fe t;
    fe_pow(r, a, &E);
    fe_sqr(&t, r);
    return fe_equal(&t, a);
}
/* secp256k1_fe_inv_var (field_5x52_impl.h:496) — value-equivalent a^(p-2) */
static void fe_inv(fe *r, const fe *a) {
    static const num E = {{0xFFFFFFFEFFFFFC2DULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
    fe_pow(r, a, &E);
}

/* ------------------------------------------------------------------ scalars mod n (scalar_4x64_impl.h) */
static const num SC_N = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
static const num SC_NHALF = {{0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL}};
static const u64 SC_NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1}; /* 2^256 - n (:23-25) */
typedef num sc;

/* secp256k1_scalar_set_b32 (:158-170): reduce mod n, report overflow */
static void sc_set_b32(sc *r, const u8 *b, int *overflow) {
    num_from_be(r, b);
    int ov = num_cmp(r, &SC_N) >= 0;
    if (ov) num_sub(r, r, &SC_N);
    if (overflow) *overflow = ov;
}
static int sc_is_high(const sc *a) { return num_cmp(a, &SC_NHALF) > 0; } /* :255-267 */
static void sc_negate(sc *r, const sc *a) { if (num_is_zero(a)) *r = *a; else num_sub(r, &SC_N, a); }
static void sc_reduce_wide(sc *r, const u64 t[8]) { /* secp256k1_scalar_reduce_512 (:384) by repeated folding */
    u64 w[8];
    memcpy(w, t, 64);
    for (int pass = 0; pass < 4; pass++) {
        u64 lo[8] = {w[0], w[1], w[2], w[3], 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) { /* lo += w[4+i] * NC << (64 i) */
            u128 c = 0;
            for (int j = 0; j < 3; j++) {
                c += (u128)w[4 + i] * SC_NC[j] + lo[i + j];
                lo[i + j] = (u64)c;
                c >>= 64;
            }
            for (int k = i + 3; k < 8 && c; k++) { c += lo[k]; lo[k] = (u64)c; c >>= 64; }
        }
        memcpy(w, lo, 64);
    }
    memcpy(r->d, w, 32);
    while (num_cmp(r, &SC_N) >= 0) num_sub(r, r, &SC_N);
}
static void sc_mul(sc *r, const sc *a, const sc *b) { u64 t[8]; mul_wide(t, a, b); sc_reduce_wide(r, t); } /* :1009 */
static void sc_inverse(sc *r, const sc *a) { /* secp256k1_scalar_inverse_var (:1139), value-equivalent a^(n-2) */
    num e = SC_N; e.d[0] -= 2;
    sc acc = {{1, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) {
        sc_mul(&acc, &acc, &acc);
        if ((e.d[i >> 6] >> (i & 63)) & 1) sc_mul(&acc, &acc, a);
    }
    *r = acc;
}

/* ------------------------------------------------------------------ group (group_impl.h) */
typedef struct { fe x, y; int inf; } ge;
typedef struct { fe x, y, z; int inf; } gej;
static const ge GE_G = {{{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
                        {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}}, 0};

static void gej_set_ge(gej *r, const ge *a) { r->x = a->x; r->y = a->y; fe_set_int(&r->z, 1); r->inf = a->inf; }
/* secp256k1_ge_is_valid_var (:356): y^2 == x^3 + 7 */
static int ge_is_valid(const ge *a) {
    fe y2, x3, seven;
    if (a->inf) return 0;
    fe_sqr(&y2, &a->y); fe_sqr(&x3, &a->x); fe_mul(&x3, &x3, &a->x);
    fe_set_int(&seven, 7); fe_add(&x3, &x3, &seven);
    return fe_equal(&y2, &x3);
}
/* secp256k1_ge_set_xo_var (:334-346) */
static int ge_set_xo(ge *r, const fe *x, int odd) {
    fe c, seven;
    fe_sqr(&c, x); fe_mul(&c, &c, x); fe_set_int(&seven, 7); fe_add(&c, &c, &seven);
    if (!fe_sqrt(&r->y, &c)) return 0;
    r->x = *x; r->inf = 0;
    if (fe_is_odd(&r->y) != odd) fe_neg(&r->y, &r->y);
    return 1;
}
/* secp256k1_gej_double_var (:474-502); formula: standard a=0 Jacobian doubling */
static void gej_double(gej *r, const gej *a) {
    if (a->inf) { *r = *a; return; }
    fe A, B, C, D, E, F, t;
    fe_sqr(&A, &a->x); fe_sqr(&B, &a->y); fe_sqr(&C, &B);
    fe_add(&t, &a->x, &B); fe_sqr(&t, &t); fe_neg(&D, &A); fe_add(&t, &t, &D); fe_neg(&D, &C); fe_add(&t, &t, &D);
    fe_add(&D, &t, &t);                      /* D = 2((X+B)^2 - A - C) */
    fe_add(&E, &A, &A); fe_add(&E, &E, &A);  /* E = 3A */
    fe_sqr(&F, &E);
    fe Z3; fe_mul(&Z3, &a->y, &a->z); fe_add(&Z3, &Z3, &Z3);
    fe X3, twoD; fe_add(&twoD, &D, &D); fe_neg(&twoD, &twoD); fe_add(&X3, &F, &twoD);
    fe Y3; fe_neg(&t, &X3); fe_add(&t, &D, &t); fe_mul(&Y3, &E, &t);
    fe c8 = C; for (int i = 0; i < 3; i++) fe_add(&c8, &c8, &c8);
    fe_neg(&c8, &c8); fe_add(&Y3, &Y3, &c8);
    r->x = X3; r->y = Y3; r->z = Z3; r->inf = 0;
}
/* secp256k1_gej_add_ge_var (:569-629) incl. the degenerate branch (:595-605) */
static void gej_add_ge(gej *r, const gej *a, const ge *b) {
    if (a->inf) { gej_set_ge(r, b); return; }
    if (b->inf) { *r = *a; return; }
    fe z12, u2, s2, h, i, t;
    fe_sqr(&z12, &a->z); fe_mul(&u2, &b->x, &z12);
    fe_mul(&s2, &b->y, &z12); fe_mul(&s2, &s2, &a->z);
    fe_neg(&t, &a->x); fe_add(&h, &u2, &t);   /* h = u2 - u1 */
    fe_neg(&t, &a->y); fe_add(&i, &s2, &t);   /* i = s2 - s1 */
    if (num_is_zero(&h)) {
        if (num_is_zero(&i)) gej_double(r, a); else { memset(r, 0, sizeof *r); r->inf = 1; }
        return;
    }
    fe h2, h3, v, X3, Y3, Z3;
    fe_sqr(&h2, &h); fe_mul(&h3, &h2, &h); fe_mul(&v, &a->x, &h2);
    fe_mul(&Z3, &a->z, &h);
    fe_sqr(&X3, &i); fe_neg(&t, &h3); fe_add(&X3, &X3, &t);
    fe_add(&t, &v, &v); fe_neg(&t, &t); fe_add(&X3, &X3, &t);
    fe_neg(&t, &X3); fe_add(&t, &v, &t); fe_mul(&Y3, &t, &i);
    fe_mul(&t, &h3, &a->y); fe_neg(&t, &t); fe_add(&Y3, &Y3, &t);
    r->x = X3; r->y = Y3; r->z = Z3; r->inf = 0;
}
/* secp256k1_ge_set_gej_var (:177) */
static void ge_set_gej(ge *r, const gej *a) {
    if (a->inf) { memset(r, 0, sizeof *r); r->inf = 1; return; }
    fe zi, zi2, zi3;
    fe_inv(&zi, &a->z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
    fe_mul(&r->x, &a->x, &zi2); fe_mul(&r->y, &a->y, &zi3); r->inf = 0;
}

/* ------------------------------------------------------------------ ecmult (ecmult_impl.h) */
#define WNAF_W 5
/* secp256k1_ecmult_wnaf (:162-218): odd digits in +-(2^(w-1)-1), >= w-1 zeros between non-zeros */
static int wnaf(int out[257], const sc *k) {
    num v = *k;
    int len = 0;
    memset(out, 0, 257 * sizeof(int));
    for (int bit = 0; bit < 257 && !num_is_zero(&v); bit++) {
        if (v.d[0] & 1) {
            int d = (int)(v.d[0] & ((1u << WNAF_W) - 1));
            if (d >= (1 << (WNAF_W - 1))) d -= (1 << WNAF_W);
            out[bit] = d;
            len = bit + 1;
            num dd = {{(u64)(d < 0 ? -d : d), 0, 0, 0}};
            if (d < 0) num_add(&v, &v, &dd); else num_sub(&v, &v, &dd);
        }
        /* v >>= 1 */
        for (int i = 0; i < 4; i++) v.d[i] = (v.d[i] >> 1) | (i < 3 ? v.d[i + 1] << 63 : 0);
    }
    return len;
}
static void odd_multiples(ge tbl[1 << (WNAF_W - 2)], const ge *a) { /* (:73-115), affine via inversion here */
    gej d, acc;
    gej_set_ge(&acc, a);
    gej_double(&d, &acc);
    ge d_aff; ge_set_gej(&d_aff, &d);
    tbl[0] = *a;
    for (int i = 1; i < (1 << (WNAF_W - 2)); i++) {
        gej_add_ge(&acc, &acc, &d_aff);
        ge_set_gej(&tbl[i], &acc);
    }
}
static ge G_TABLE[1 << (WNAF_W - 2)];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void g_init(void) { odd_multiples(G_TABLE, &GE_G); }

/* R = na*A + ng*G : secp256k1_ecmult (:346) / secp256k1_ecmult_strauss_wnaf (:234-341) */
static void ecmult(gej *r, const ge *a, const sc *na, const sc *ng) {
    int wa[257], wg[257];
    ge ta[1 << (WNAF_W - 2)];
    pthread_once(&g_once, g_init);
    int la = wnaf(wa, na), lg = wnaf(wg, ng);
    if (la) odd_multiples(ta, a);
    int len = la > lg ? la : lg;
    memset(r, 0, sizeof *r); r->inf = 1;
    for (int i = len - 1; i >= 0; i--) {
        gej_double(r, r);
        int d;
        if ((d = wa[i]) != 0) {
            ge t = ta[(d < 0 ? -d : d) >> 1];
            if (d < 0) fe_neg(&t.y, &t.y);
            gej_add_ge(r, r, &t);
        }
        if ((d = wg[i]) != 0) {
            ge t = G_TABLE[(d < 0 ? -d : d) >> 1];
            if (d < 0) fe_neg(&t.y, &t.y);
            gej_add_ge(r, r, &t);
        }
    }
}

/* ------------------------------------------------------------------ SHA-256 (ccan/ccan/crypto/sha256/sha256.c:87,243; hash_impl.h) */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
typedef struct { uint32_t h[8]; u8 buf[64]; u64 len; } sha_ctx;
static uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(uint32_t h[8], const u8 *p) {
    uint32_t w[64], a[8];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++)
        w[i] = w[i - 16] + (ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] +
               (ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10));
    memcpy(a, h, 32);
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = a[7] + (ror(a[4], 6) ^ ror(a[4], 11) ^ ror(a[4], 25)) + ((a[4] & a[5]) ^ (~a[4] & a[6])) + K256[i] + w[i];
        uint32_t t2 = (ror(a[0], 2) ^ ror(a[0], 13) ^ ror(a[0], 22)) + ((a[0] & a[1]) ^ (a[0] & a[2]) ^ (a[1] & a[2]));
        memmove(a + 1, a, 28);
        a[4] += t1;
        a[0] = t1 + t2;
    }
    for (int i = 0; i < 8; i++) h[i] += a[i];
}
static void sha_init(sha_ctx *c) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c->h, iv, 32); c->len = 0;
}
static void sha_write(sha_ctx *c, const u8 *p, size_t n) {
    while (n) {
        size_t off = c->len & 63, take = 64 - off;
        if (take > n) take = n;
        memcpy(c->buf + off, p, take);
        c->len += take; p += take; n -= take;
        if ((c->len & 63) == 0) sha_block(c->h, c->buf);
    }
}
static void sha_final(sha_ctx *c, u8 out[32]) {
    u64 bits = c->len * 8;
    u8 pad[72] = {0x80};
    size_t padlen = 1 + ((119 - (c->len & 63)) & 63);
    sha_write(c, pad, padlen);
    u8 lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (u8)(bits >> (56 - 8 * i));
    sha_write(c, lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = (u8)(c->h[i] >> 24); out[4 * i + 1] = (u8)(c->h[i] >> 16); out[4 * i + 2] = (u8)(c->h[i] >> 8); out[4 * i + 3] = (u8)c->h[i]; }
}
void port_sha256(const u8 *p, size_t n, u8 out[32]) { sha_ctx c; sha_init(&c); sha_write(&c, p, n); sha_final(&c, out); }
/* sha256_double (/root/reference/bitcoin/shadouble.c:7-11) */
void port_sha256d(const u8 *p, size_t n, u8 out[32]) { u8 t[32]; port_sha256(p, n, t); port_sha256(t, 32, out); }

/* ------------------------------------------------------------------ ECDSA (secp256k1.c, ecdsa_impl.h, eckey_impl.h) */
/* secp256k1_eckey_pubkey_parse (eckey_impl.h:17-35) for the 33- and 65-byte encodings */
static int pubkey_parse(ge *q, const u8 *pub, size_t len) {
    fe x, y;
    if (len == 33 && (pub[0] == 2 || pub[0] == 3)) {
        return fe_set_b32_limit(&x, pub + 1) && ge_set_xo(q, &x, pub[0] == 3);
    } else if (len == 65 && (pub[0] == 4 || pub[0] == 6 || pub[0] == 7)) {
        if (!fe_set_b32_limit(&x, pub + 1) || !fe_set_b32_limit(&y, pub + 33)) return 0;
        q->x = x; q->y = y; q->inf = 0;
        if ((pub[0] == 6 || pub[0] == 7) && fe_is_odd(&y) != (pub[0] == 7)) return 0;
        return ge_is_valid(q);
    }
    return 0;
}
/* secp256k1_ecdsa_signature_parse_compact (secp256k1.c:377-396) + secp256k1_ecdsa_verify (:442-456)
 * + secp256k1_ecdsa_sig_verify (ecdsa_impl.h:195-264) */
static int ecdsa_verify(const u8 *msg32, const ge *q, const u8 *sig64) {
    sc r, s, m, sn, u1, u2;
    int ov;
    sc_set_b32(&r, sig64, &ov); if (ov) return 0;          /* parse_compact: r >= n */
    sc_set_b32(&s, sig64 + 32, &ov); if (ov) return 0;     /* parse_compact: s >= n */
    sc_set_b32(&m, msg32, NULL);                           /* message reduced, never rejected */
    if (sc_is_high(&s)) return 0;                          /* secp256k1.c:451 low-S rule */
    if (num_is_zero(&r) || num_is_zero(&s)) return 0;      /* ecdsa_impl.h:204 */
    sc_inverse(&sn, &s); sc_mul(&u1, &sn, &m); sc_mul(&u2, &sn, &r);
    gej R; ecmult(&R, q, &u2, &u1);
    if (R.inf) return 0;                                   /* :213 */
    /* secp256k1_gej_eq_x_var: r*Z^2 == X, then the r+n candidate if r < p-n (:229-264) */
    fe xr = r, zz, t;
    fe_sqr(&zz, &R.z); fe_mul(&t, &xr, &zz);
    if (fe_equal(&t, &R.x)) return 1;
    static const num P_MINUS_N = {{0x402DA1722FC9BAEEULL, 0x4551231950B75FC4ULL, 1, 0}};
    if (num_cmp(&r, &P_MINUS_N) >= 0) return 0;
    num_add(&xr, &r, &SC_N);
    fe_mul(&t, &xr, &zz);
    return fe_equal(&t, &R.x);
}

/* ------------------------------------------------------------------ BIP-340 (modules/schnorrsig/main_impl.h, modules/extrakeys/main_impl.h) */
static int schnorr_verify(const u8 *msg32, const u8 *xonly32, const u8 *sig64) {
    fe px, rx;
    ge P;
    sc s, e;
    int ov;
    if (!fe_set_b32_limit(&px, xonly32)) return 0;          /* extrakeys/main_impl.h:32 */
    if (!ge_set_xo(&P, &px, 0)) return 0;                   /* :35 */
    if (!fe_set_b32_limit(&rx, sig64)) return 0;            /* schnorrsig/main_impl.h:235 */
    sc_set_b32(&s, sig64 + 32, &ov); if (ov) return 0;      /* :239-242 */
    /* secp256k1_schnorrsig_challenge (:116-127): tagged hash of r || P.x || msg */
    u8 tag[32], h[32];
    port_sha256((const u8 *)"BIP0340/challenge", 17, tag);
    sha_ctx c; sha_init(&c);
    sha_write(&c, tag, 32); sha_write(&c, tag, 32);
    sha_write(&c, sig64, 32); sha_write(&c, xonly32, 32); sha_write(&c, msg32, 32);
    sha_final(&c, h);
    sc_set_b32(&e, h, NULL);
    sc_negate(&e, &e);                                      /* :250 */
    gej Rj; ecmult(&Rj, &P, &e, &s);                        /* :251 */
    ge R; ge_set_gej(&R, &Rj);
    if (R.inf) return 0;                                    /* :254 */
    if (fe_is_odd(&R.y)) return 0;                          /* :259 */
    return fe_equal(&rx, &R.x);                             /* :262 */
}

/* ------------------------------------------------------------------ batch entry points (same shapes as oracle/ref_harness.c) */
typedef struct { int kind; const u8 *msg, *pub, *sig; size_t lo, hi; u8 *out; } job_t;
static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        ge q;
        if (j->kind == 0) {
            j->out[i] = (u8)(pubkey_parse(&q, j->pub + 33 * i, 33) && ecdsa_verify(j->msg + 32 * i, &q, j->sig + 64 * i));
        } else if (j->kind == 1) {
            u8 pk[65]; pk[0] = 4; memcpy(pk + 1, j->pub + 64 * i, 64);
            j->out[i] = (u8)(pubkey_parse(&q, pk, 65) && ecdsa_verify(j->msg + 32 * i, &q, j->sig + 64 * i));
        } else {
            j->out[i] = (u8)schnorr_verify(j->msg + 32 * i, j->pub + 32 * i, j->sig + 64 * i);
        }
    }
    return NULL;
}
static void run(int kind, const u8 *msg, const u8 *pub, const u8 *sig, size_t n, u8 *out, int nthreads) {
    pthread_once(&g_once, g_init);
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n && n) nthreads = (int)n;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t *jobs = malloc(sizeof(job_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){kind, msg, pub, sig, n * (size_t)t / (size_t)nthreads, n * (size_t)(t + 1) / (size_t)nthreads, out};
        if (nthreads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
void port_ecdsa_verify_batch(const u8 *msg, const u8 *pub33, const u8 *sig, size_t n, u8 *out, int nthreads) { run(0, msg, pub33, sig, n, out, nthreads); }
void port_ecdsa_verify_batch_xy(const u8 *msg, const u8 *pubxy64, const u8 *sig, size_t n, u8 *out, int nthreads) { run(1, msg, pubxy64, sig, n, out, nthreads); }
void port_schnorr_verify_batch(const u8 *msg, const u8 *xonly32, const u8 *sig, size_t n, u8 *out, int nthreads) { run(2, msg, xonly32, sig, n, out, nthreads); }
/* x*G, uncompressed without prefix (ecmult KAT support) ; returns 0 for the point at infinity */
int port_scalar_base_mult(const u8 *scalar32, u8 *xy64) {
    sc k, zero = {{0, 0, 0, 0}};
    sc_set_b32(&k, scalar32, NULL);
    gej R; ge A;
    ecmult(&R, &GE_G, &zero, &k);
    ge_set_gej(&A, &R);
    if (A.inf) return 0;
    fe_get_b32(xy64, &A.x); fe_get_b32(xy64 + 32, &A.y);
    return 1;
}
int port_pubkey_parse33(const u8 *pub33, u8 *xy64) {
    ge q;
    if (!pubkey_parse(&q, pub33, 33)) return 0;
    fe_get_b32(xy64, &q.x); fe_get_b32(xy64 + 32, &q.y);
    return 1;
}
