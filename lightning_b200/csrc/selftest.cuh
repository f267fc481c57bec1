// selftest.cuh — ONE primitive of the engine's arithmetic on caller operands (test support).
//
// Shared by the k_selftest kernel (engine.cu; the inline-PTX device forms, reached through sv_selftest_host) and by
// tests/host_emul (the uint64 host forms), so that the same test code runs against both.  Operands A, B are 8
// little-endian 32-bit limbs each, R receives 16 limbs (see the SV_ST_* list in include/cln_sigverify.h).
// Model: libsecp256k1 tests.c:3023-3176 (field self-tests), :2354 (scalar tests).
#pragma once
#include "../../include/cln_sigverify.h"
#include "verify.cuh"

SV_HD void selftest_item(int op, const u32 A[8], const u32 B[8], u32 R[16], const ge_mem* gtab) {
    SV_UNROLL
    for (int k = 0; k < 16; k++) R[k] = 0;
    fe x, y, r;
    sc p, q, s;
SV_UNROLL
    for (int k = 0; k < 8; k++) { x.v[k] = A[k]; y.v[k] = B[k]; p.v[k] = A[k]; q.v[k] = B[k]; }
    fe_set_zero(r);
    bool is_fe = false, is_sc = false;
    switch (op) {
        case SV_ST_FE_MUL: fe_mul(r, x, y); is_fe = true; break;
        case SV_ST_FE_SQR: fe_sqr(r, x); is_fe = true; break;
        case SV_ST_FE_ADD: fe_add(r, x, y); is_fe = true; break;
        case SV_ST_FE_SUB: fe_sub(r, x, y); is_fe = true; break;
        case SV_ST_FE_NEG: fe_neg(r, x); is_fe = true; break;
        case SV_ST_FE_NORMALIZE: r = x; fe_normalize(r); is_fe = true; R[8] = fe_is_zero(x); R[9] = fe_gte_p(x); R[10] = fe_equal(x, y); break;
        case SV_ST_FE_INV: fe_inv(r, x); is_fe = true; break;
        case SV_ST_FE_SQRT: R[8] = fe_sqrt(r, x) ? 1u : 0u; is_fe = true; break;
        case SV_ST_FE_MUL3: fe_mul3(r, x); is_fe = true; break;
        case SV_ST_FE_MUL8: fe_mul8(r, x); is_fe = true; break;
        case SV_ST_FE_MUL_SMALL: fe_mul_small(r, x, B[0] & 0xFFFFu); is_fe = true; break;
        case SV_ST_FE_DBL: fe_dbl(r, x); is_fe = true; break;
        case SV_ST_FE_B32: {  // set_b32 / get_b32 round trip of the big-endian bytes held in A (as memory order)
            u8 bytes[32], back[32];
SV_UNROLL
            for (int k = 0; k < 8; k++) { bytes[4 * k] = (u8)A[k]; bytes[4 * k + 1] = (u8)(A[k] >> 8); bytes[4 * k + 2] = (u8)(A[k] >> 16); bytes[4 * k + 3] = (u8)(A[k] >> 24); }
            R[8] = fe_set_b32(r, bytes) ? 1u : 0u;
            fe_get_b32(back, r);
SV_UNROLL
            for (int k = 0; k < 8; k++) R[k] = (u32)back[4 * k] | ((u32)back[4 * k + 1] << 8) | ((u32)back[4 * k + 2] << 16) | ((u32)back[4 * k + 3] << 24);
            break;
        }
        case SV_ST_U256_MUL_WIDE: u256_mul_wide(R, A, B); break;
        case SV_ST_U256_SQR_WIDE: u256_sqr_wide(R, A); break;
        case SV_ST_FE_REDUCE512: {
            u32 t[16];
SV_UNROLL
            for (int k = 0; k < 8; k++) { t[k] = A[k]; t[8 + k] = B[k]; }
            fe_reduce512(r, t);
            is_fe = true;
            break;
        }
        case SV_ST_U256_ADD: R[8] = u256_add(R, A, B); break;
        case SV_ST_U256_SUB: R[8] = u256_sub(R, A, B); break;
        case SV_ST_SC_MUL: sc_mul(s, p, q); is_sc = true; break;
        case SV_ST_SC_SQR: sc_sqr(s, p); is_sc = true; break;
        case SV_ST_SC_ADD: sc_add(s, p, q); is_sc = true; break;
        case SV_ST_SC_NEGATE: sc_negate(s, p); is_sc = true; R[8] = sc_is_high(p); R[9] = sc_is_zero(p); R[10] = sc_gte_n(A); break;
        case SV_ST_SC_INVERSE: sc_inverse(s, p); is_sc = true; break;
        case SV_ST_SC_INVERSE_VAR: sc_inverse_var(s, p); is_sc = true; break;
        case SV_ST_FE_INV_VAR: fe_inv_var(r, x); is_fe = true; break;
        case SV_ST_SC_REDUCE512: {
            u32 t[16];
SV_UNROLL
            for (int k = 0; k < 8; k++) { t[k] = A[k]; t[8 + k] = B[k]; }
            sc_reduce512(s, t);
            is_sc = true;
            break;
        }
        case SV_ST_SC_SPLIT_LAMBDA: {
            sc r1, r2;
            sc_split_lambda(r1, r2, p);
SV_UNROLL
            for (int k = 0; k < 8; k++) { R[k] = r1.v[k]; R[8 + k] = r2.v[k]; }
            break;
        }
        case SV_ST_SC_SET_B32: {  // scalar_set_b32 of the 32 bytes held in A (memory order): reduced value + overflow flag
            u8 bytes[32];
            bool ov;
SV_UNROLL
            for (int k = 0; k < 8; k++) { bytes[4 * k] = (u8)A[k]; bytes[4 * k + 1] = (u8)(A[k] >> 8); bytes[4 * k + 2] = (u8)(A[k] >> 16); bytes[4 * k + 3] = (u8)(A[k] >> 24); }
            sc_set_b32(s, bytes, &ov);
            is_sc = true;
            R[8] = ov ? 1u : 0u;
            break;
        }
        case SV_ST_ECMULT_GEN: {  // A = scalar k (< n): affine k*G through the fixed-base comb table (x -> R[0..7], y -> R[8..15])
            ge P;
            if (sc_is_zero(p)) break;  // infinity: all-zero output
            ecmult_gen_comb(P, p, gtab);
SV_UNROLL
            for (int k = 0; k < 8; k++) { R[k] = P.x.v[k]; R[8 + k] = P.y.v[k]; }
            break;
        }
        case SV_ST_PREPARE_U2: {  // B = u2: GLV split forced odd, sign/magnitude halves as the work record stores them
            sv_work w;
            sc_prepare_u2(w, q);
SV_UNROLL
            for (int k = 0; k < 5; k++) { R[k] = w.k1[k]; R[5 + k] = w.k2[k]; }
            break;
        }
        case SV_ST_PREPARE_U1: {  // A = u1: the 16 signed comb digits
            sv_work w;
            sc_prepare_u1(w, p);
SV_UNROLL
            for (int k = 0; k < 16; k++) R[k] = (u32)w.gd[k];
            break;
        }
        default: break;
    }
    if (is_fe) {
SV_UNROLL
        for (int k = 0; k < 8; k++) R[k] = r.v[k];
    }
    if (is_sc) {
SV_UNROLL
        for (int k = 0; k < 8; k++) R[k] = s.v[k];
    }
}
