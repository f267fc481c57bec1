/*
 * oracle/ref_harness.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin batch harness over the UNMODIFIED reference implementation (libsecp256k1-zkp as
 * vendored by the reference at external/libwally-core/src/secp256k1/, and CCAN sha256 at
 * ccan/ccan/crypto/sha256/sha256.c).  The reference sources are compiled where they lie
 * under /root/reference by oracle/Makefile; only this file (our own code, public API calls
 * only) lives in the repo.  Output: oracle/_ref/libsecp_ref.so.
 *
 * What each entry point times/checks is exactly the per-item sequence BASELINE.md §3 names:
 *   ECDSA   : secp256k1_ec_pubkey_parse + secp256k1_ecdsa_signature_parse_compact +
 *             secp256k1_ecdsa_verify            (reference: secp256k1.c:270,377,442)
 *   Schnorr : secp256k1_xonly_pubkey_parse + secp256k1_schnorrsig_verify(msglen 32)
 *             (reference: modules/extrakeys/main_impl.h:23, modules/schnorrsig/main_impl.h:219)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <secp256k1.h>
#include <secp256k1_extrakeys.h>
#include <secp256k1_schnorrsig.h>

static secp256k1_context *g_ctx;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void ctx_init(void) { g_ctx = secp256k1_context_create(SECP256K1_CONTEXT_NONE); }
static secp256k1_context *ctx(void) { pthread_once(&g_once, ctx_init); return g_ctx; }

typedef struct {
    int kind; /* 0 ecdsa33, 1 ecdsa65(04|x|y given as 64 raw bytes), 2 schnorr */
    const uint8_t *msg, *pub, *sig;
    size_t lo, hi;
    uint8_t *out;
} job_t;

static int verify_ecdsa(const uint8_t *msg32, const uint8_t *pub, size_t publen, const uint8_t *sig64) {
    secp256k1_pubkey pk;
    secp256k1_ecdsa_signature s;
    if (!secp256k1_ec_pubkey_parse(ctx(), &pk, pub, publen)) return 0;
    if (!secp256k1_ecdsa_signature_parse_compact(ctx(), &s, sig64)) return 0;
    return secp256k1_ecdsa_verify(ctx(), &s, msg32, &pk);
}

static int verify_schnorr(const uint8_t *msg32, const uint8_t *xonly32, const uint8_t *sig64) {
    secp256k1_xonly_pubkey pk;
    if (!secp256k1_xonly_pubkey_parse(ctx(), &pk, xonly32)) return 0;
    return secp256k1_schnorrsig_verify(ctx(), sig64, msg32, 32, &pk);
}

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        if (j->kind == 0) {
            j->out[i] = (uint8_t)verify_ecdsa(j->msg + 32 * i, j->pub + 33 * i, 33, j->sig + 64 * i);
        } else if (j->kind == 1) {
            uint8_t pk65[65];
            pk65[0] = 4;
            memcpy(pk65 + 1, j->pub + 64 * i, 64);
            j->out[i] = (uint8_t)verify_ecdsa(j->msg + 32 * i, pk65, 65, j->sig + 64 * i);
        } else {
            j->out[i] = (uint8_t)verify_schnorr(j->msg + 32 * i, j->pub + 32 * i, j->sig + 64 * i);
        }
    }
    return NULL;
}

static void run(int kind, const uint8_t *msg, const uint8_t *pub, const uint8_t *sig, size_t n,
                uint8_t *out, int nthreads) {
    (void)ctx();
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n && n > 0) nthreads = (int)n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){kind, msg, pub, sig, n * (size_t)t / (size_t)nthreads,
                          n * (size_t)(t + 1) / (size_t)nthreads, out};
        if (nthreads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* SoA inputs: msg[n][32], pub[n][33|64|32], sig[n][64]; out[n] = 0/1. */
void ref_ecdsa_verify_batch(const uint8_t *msg, const uint8_t *pub33, const uint8_t *sig, size_t n,
                            uint8_t *out, int nthreads) { run(0, msg, pub33, sig, n, out, nthreads); }
void ref_ecdsa_verify_batch_xy(const uint8_t *msg, const uint8_t *pubxy64, const uint8_t *sig, size_t n,
                               uint8_t *out, int nthreads) { run(1, msg, pubxy64, sig, n, out, nthreads); }
void ref_schnorr_verify_batch(const uint8_t *msg, const uint8_t *xonly32, const uint8_t *sig, size_t n,
                              uint8_t *out, int nthreads) { run(2, msg, xonly32, sig, n, out, nthreads); }

/* ---- helpers for fixture generation (tests only) ---- */
int ref_ecdsa_sign(const uint8_t *seckey32, const uint8_t *msg32, uint8_t *sig64_out) {
    secp256k1_ecdsa_signature s;
    if (!secp256k1_ecdsa_sign(ctx(), &s, msg32, seckey32, NULL, NULL)) return 0;
    return secp256k1_ecdsa_signature_serialize_compact(ctx(), sig64_out, &s);
}
int ref_pubkey_create(const uint8_t *seckey32, uint8_t *pub33_out, uint8_t *pubxy64_out) {
    secp256k1_pubkey pk;
    uint8_t u[65];
    size_t l = 33;
    if (!secp256k1_ec_pubkey_create(ctx(), &pk, seckey32)) return 0;
    secp256k1_ec_pubkey_serialize(ctx(), pub33_out, &l, &pk, SECP256K1_EC_COMPRESSED);
    l = 65;
    secp256k1_ec_pubkey_serialize(ctx(), u, &l, &pk, SECP256K1_EC_UNCOMPRESSED);
    if (pubxy64_out) memcpy(pubxy64_out, u + 1, 64);
    return 1;
}
int ref_schnorr_sign(const uint8_t *seckey32, const uint8_t *msg32, uint8_t *sig64_out, uint8_t *xonly32_out) {
    secp256k1_keypair kp;
    secp256k1_xonly_pubkey xo;
    if (!secp256k1_keypair_create(ctx(), &kp, seckey32)) return 0;
    if (!secp256k1_schnorrsig_sign32(ctx(), sig64_out, msg32, &kp, NULL)) return 0;
    secp256k1_keypair_xonly_pub(ctx(), &xo, NULL, &kp);
    return secp256k1_xonly_pubkey_serialize(ctx(), xonly32_out, &xo);
}
/* DER -> compact (returns 0 if the reference's strict DER parser rejects it). */
int ref_sig_der_to_compact(const uint8_t *der, size_t len, uint8_t *sig64_out) {
    secp256k1_ecdsa_signature s;
    if (!secp256k1_ecdsa_signature_parse_der(ctx(), &s, der, len)) return 0;
    return secp256k1_ecdsa_signature_serialize_compact(ctx(), sig64_out, &s);
}
/* any SEC1 pubkey encoding -> 33-byte compressed + raw x|y */
int ref_pubkey_convert(const uint8_t *in, size_t inlen, uint8_t *pub33_out, uint8_t *pubxy64_out) {
    secp256k1_pubkey pk;
    uint8_t u[65];
    size_t l = 33;
    if (!secp256k1_ec_pubkey_parse(ctx(), &pk, in, inlen)) return 0;
    secp256k1_ec_pubkey_serialize(ctx(), pub33_out, &l, &pk, SECP256K1_EC_COMPRESSED);
    l = 65;
    secp256k1_ec_pubkey_serialize(ctx(), u, &l, &pk, SECP256K1_EC_UNCOMPRESSED);
    memcpy(pubxy64_out, u + 1, 64);
    return 1;
}
/* x*G serialised uncompressed-without-prefix (for the ecmult KAT and G-table checks) */
int ref_scalar_base_mult(const uint8_t *scalar32, uint8_t *xy64_out) {
    return ref_pubkey_create(scalar32, (uint8_t[33]){0}, xy64_out);
}

/* CCAN sha256 (reference: ccan/ccan/crypto/sha256/sha256.c:243) and bitcoin/shadouble.c:7 semantics */
struct sha256 { union { uint32_t u32[8]; unsigned char u8[32]; } u; };
void sha256(struct sha256 *sha, const void *p, size_t size);
void ref_sha256(const uint8_t *p, size_t len, uint8_t *out32) {
    struct sha256 h; sha256(&h, p, len); memcpy(out32, h.u.u8, 32);
}
void ref_sha256d(const uint8_t *p, size_t len, uint8_t *out32) {
    struct sha256 h, h2; sha256(&h, p, len); sha256(&h2, &h, sizeof(h)); memcpy(out32, h2.u.u8, 32);
}
/* tagged hash as used by the reference (hash_impl.h secp256k1_sha256_initialize_tagged) */
void ref_tagged_sha256(const uint8_t *tag, size_t taglen, const uint8_t *msg, size_t msglen, uint8_t *out32) {
    secp256k1_tagged_sha256(ctx(), out32, tag, taglen, msg, msglen);
}

/* ---- seeded workload for bench.py's reference arm (SURVEY.md §8(d) C2 recipe): SplitMix64 secret
 * keys and message hashes, signed with the reference's RFC6979 signer (low-S by construction); every
 * 10th item corrupted, classes round-robin {msg bit, r bit, s bit, high-S, neighbour's key,
 * non-residue x, prefix 0x04}.  Generation is untimed. ---- */
static uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t seed; size_t lo, hi; uint8_t *msg, *pub, *sig; } gen_t;
static void *gen_worker(void *arg) {
    gen_t *g = (gen_t *)arg;
    for (size_t i = g->lo; i < g->hi; i++) {
        uint64_t st = g->seed ^ (0xD1B54A32D192ED03ULL * (i + 1));
        uint8_t sk[32];
        for (int k = 0; k < 4; k++) { uint64_t v = splitmix(&st); memcpy(sk + 8 * k, &v, 8); }
        for (int k = 0; k < 4; k++) { uint64_t v = splitmix(&st); memcpy(g->msg + 32 * i + 8 * k, &v, 8); }
        sk[0] &= 0x7f; sk[31] |= 1; /* 0 < sk < n */
        ref_pubkey_create(sk, g->pub + 33 * i, NULL);
        ref_ecdsa_sign(sk, g->msg + 32 * i, g->sig + 64 * i);
    }
    return NULL;
}
void ref_make_ecdsa_batch(uint64_t seed, size_t n, uint8_t *msg, uint8_t *pub33, uint8_t *sig, int nthreads) {
    (void)ctx();
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    gen_t *jobs = (gen_t *)malloc(sizeof(gen_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (gen_t){seed, n * (size_t)t / (size_t)nthreads, n * (size_t)(t + 1) / (size_t)nthreads, msg, pub33, sig};
        pthread_create(&th[t], NULL, gen_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    static const uint8_t order[32] = {0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFE,
                                      0xBA,0xAE,0xDC,0xE6,0xAF,0x48,0xA0,0x3B,0xBF,0xD2,0x5E,0x8C,0xD0,0x36,0x41,0x41};
    for (size_t i = 0, cls = 0; i < n; i += 10, cls++) {
        size_t j = (i + 1) % n;
        switch (cls % 7) {
        case 0: msg[32 * i + 5] ^= 4; break;
        case 1: sig[64 * i + 7] ^= 1; break;
        case 2: sig[64 * i + 40] ^= 1; break;
        case 3: { /* s <- n - s */
            int borrow = 0;
            for (int k = 31; k >= 0; k--) {
                int d = (int)order[k] - (int)sig[64 * i + 32 + k] - borrow;
                borrow = d < 0; sig[64 * i + 32 + k] = (uint8_t)(d + (borrow << 8));
            }
            break; }
        case 4: memcpy(pub33 + 33 * i, pub33 + 33 * j, 33); break;
        case 5: memset(pub33 + 33 * i + 1, 0, 32); pub33[33 * i + 32] = 5; break; /* x = 5: 5^3+7 is a non-residue */
        case 6: pub33[33 * i] = 4; break;
        }
    }
}

/* BIP-340 counterpart (config C3): x-only keys, secp256k1_schnorrsig_sign32 with aux_rand = NULL (deterministic); every
 * 10th item corrupted, classes round-robin {msg bit, R.x bit, s bit, r >= p, s >= n, negated R (odd y), neighbour's key,
 * key x not on the curve}. */
typedef struct { uint64_t seed; size_t lo, hi; uint8_t *msg, *xonly, *sig; } sgen_t;
static void *sgen_worker(void *arg) {
    sgen_t *g = (sgen_t *)arg;
    for (size_t i = g->lo; i < g->hi; i++) {
        uint64_t st = g->seed ^ (0xD1B54A32D192ED03ULL * (i + 1));
        uint8_t sk[32];
        for (int k = 0; k < 4; k++) { uint64_t v = splitmix(&st); memcpy(sk + 8 * k, &v, 8); }
        for (int k = 0; k < 4; k++) { uint64_t v = splitmix(&st); memcpy(g->msg + 32 * i + 8 * k, &v, 8); }
        sk[0] &= 0x7f; sk[31] |= 1;
        ref_schnorr_sign(sk, g->msg + 32 * i, g->sig + 64 * i, g->xonly + 32 * i);
    }
    return NULL;
}
void ref_make_schnorr_batch(uint64_t seed, size_t n, uint8_t *msg, uint8_t *xonly32, uint8_t *sig, int nthreads) {
    (void)ctx();
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    sgen_t *jobs = (sgen_t *)malloc(sizeof(sgen_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (sgen_t){seed, n * (size_t)t / (size_t)nthreads, n * (size_t)(t + 1) / (size_t)nthreads, msg, xonly32, sig};
        pthread_create(&th[t], NULL, sgen_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    static const uint8_t fieldp[32] = {0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,
                                       0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFE,0xFF,0xFF,0xFC,0x2F};
    for (size_t i = 0, cls = 0; i < n; i += 10, cls++) {
        size_t j = (i + 1) % n;
        switch (cls % 8) {
        case 0: msg[32 * i + 5] ^= 4; break;
        case 1: sig[64 * i + 7] ^= 1; break;
        case 2: sig[64 * i + 40] ^= 1; break;
        case 3: memset(sig + 64 * i, 0xFF, 32); break;      /* r >= p */
        case 4: memset(sig + 64 * i + 32, 0xFF, 32); break; /* s >= n */
        case 5: { /* R.x kept, so R = lift_x is unchanged; instead negate s: verifies to -R side -> x differs */
            int borrow = 0;
            static const uint8_t order[32] = {0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFF,0xFE,
                                              0xBA,0xAE,0xDC,0xE6,0xAF,0x48,0xA0,0x3B,0xBF,0xD2,0x5E,0x8C,0xD0,0x36,0x41,0x41};
            for (int k = 31; k >= 0; k--) {
                int d = (int)order[k] - (int)sig[64 * i + 32 + k] - borrow;
                borrow = d < 0; sig[64 * i + 32 + k] = (uint8_t)(d + (borrow << 8));
            }
            break; }
        case 6: memcpy(xonly32 + 32 * i, xonly32 + 32 * j, 32); break;
        case 7: memset(xonly32 + 32 * i, 0, 32); xonly32[32 * i + 31] = 5; break; /* x = 5 is not on the curve */
        }
        (void)fieldp;
    }
}

/* ---- opaque libsecp256k1 structs for the drop-in tests (what CLN's wire parsers hand to check_signed_hash) ---- */
int ref_make_opaque_pubkey(const uint8_t *pub33, uint8_t *opaque64) {
    secp256k1_pubkey pk;
    if (!secp256k1_ec_pubkey_parse(ctx(), &pk, pub33, 33)) return 0;
    memcpy(opaque64, pk.data, 64);
    return 1;
}
int ref_make_opaque_sig(const uint8_t *sig64, uint8_t *opaque64) {
    secp256k1_ecdsa_signature s;
    if (!secp256k1_ecdsa_signature_parse_compact(ctx(), &s, sig64)) return 0;
    memcpy(opaque64, s.data, 64);
    return 1;
}
/* reference verdicts on opaque inputs: secp256k1_ecdsa_verify as bitcoin/signature.c:174 calls it, and
 * check_schnorr_sig's serialize -> drop parity -> xonly_parse -> schnorrsig_verify (signature.c:408-430) */
int ref_check_signed_hash_opaque(const uint8_t *hash32, const uint8_t *sig_opaque64, const uint8_t *pub_opaque64) {
    secp256k1_pubkey pk; secp256k1_ecdsa_signature s;
    memcpy(pk.data, pub_opaque64, 64); memcpy(s.data, sig_opaque64, 64);
    return secp256k1_ecdsa_verify(ctx(), &s, hash32, &pk);
}
int ref_check_schnorr_sig_opaque(const uint8_t *hash32, const uint8_t *pub_opaque64, const uint8_t *sig64) {
    secp256k1_pubkey pk; secp256k1_xonly_pubkey xo; uint8_t ser[33]; size_t l = 33;
    memcpy(pk.data, pub_opaque64, 64);
    if (!secp256k1_ec_pubkey_serialize(ctx(), ser, &l, &pk, SECP256K1_EC_COMPRESSED)) return -1;
    if (!secp256k1_xonly_pubkey_parse(ctx(), &xo, ser + 1)) return -1;
    return secp256k1_schnorrsig_verify(ctx(), sig64, hash32, 32, &xo);
}
