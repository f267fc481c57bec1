// batch.cuh — BIP-340 batch verification by random linear combination.
//
// n signatures (r_i, s_i) on messages m_i under x-only keys P_i are all valid iff (with overwhelming probability over the
// random a_i, BIP-340 "Batch Verification")
//        sum_i a_i*R_i  +  sum_i (a_i*e_i)*P_i  -  (sum_i a_i*s_i)*G  ==  infinity ,     R_i = lift_x(r_i), P_i = lift_x(px_i)
// The reference ships the multi-scalar machinery (secp256k1_ecmult_multi_var / Pippenger, ecmult_impl.h:50-56, ecmult.h:49-59)
// but no batch-verify API; this is our own schedule for it, built for a GPU:
//
//   * the batch is cut into GROUPS of SV_SB_GROUP signatures; each group gets its own equation, so one bad signature only
//     sends its group (not the batch) back to one-by-one verification — verdicts stay exact per signature;
//   * a_i = alpha_i + beta_i*lambda with 64-bit alpha_i (odd), beta_i drawn from SHA-256(seed || i): 2^127 equally likely
//     values per signature, and the R-terms become two 64-bit scalar multiplications alpha_i*R_i + beta_i*(lambda R_i);
//     c_i = a_i*e_i mod n is GLV-split into two <= 129-bit halves, as in single verification;
//   * all scalars are recoded into signed 6-bit digits; per (group, window) ONE WARP runs the bucket method: lane b owns
//     bucket |digit| = b+1, collects its points (counting sort in shared memory), adds them with mixed additions, and the
//     32 bucket sums are folded into sum_b (b+1)*B_b by a parallel suffix scan + tree reduction in shared memory;
//   * one thread per group combines the window sums (Horner, 6 doublings per window), subtracts (sum a_i s_i)*G through the
//     fixed-base comb and tests for infinity.
// A signature whose encoding already fails (r >= p, s >= n, r or px not an x coordinate) is excluded from its group's
// equation and gets verdict 0 at once (main_impl.h:235-242, extrakeys/main_impl.h:32-38).
#pragma once
#include "verify.cuh"

#define SV_SB_GROUP 1024    // signatures per equation
#define SV_SB_WINDOWS 23    // 23 x 6 bits = 138 >= 129-bit halves + recoding carry
#define SV_SB_TERMS 4       // alpha*R, beta*lambdaR, k1*P, k2*lambdaP

// signed base-64 digits of a sign-magnitude value (5 limbs, < 2^131): dig[w*stride], each in [-32, 32]
SV_HD void sb_recode(signed char* dig, size_t stride, const u32 mag[5], u32 neg) {
    u32 carry = 0;
    for (int w = 0; w < SV_SB_WINDOWS; w++) {
        int off = 6 * w;
        int l = off >> 5, sh = off & 31;
        u64 two = (l < 5 ? (u64)mag[l] : 0) | ((l + 1 < 5) ? ((u64)mag[l + 1] << 32) : 0);
        u32 d = ((u32)(two >> sh) & 63u) + carry;
        int v;
        if (d > 32u) { v = (int)d - 64; carry = 1; } else { v = (int)d; carry = 0; }
        dig[(size_t)w * stride] = (signed char)(neg ? -v : v);
    }
}

// per-signature preparation.  pts: 2 entries (R_i, P_i: x, y, beta*x), dig: 4 digit columns, t = a_i*s_i.  Returns the
// encoding check; on failure all digits are 0 and t = 0, so the item drops out of the equation.
SV_HD bool sb_prepare(const u8* msg32, const u8* xonly32, const u8* sig64, const u8* seed32, u64 index, qtab_entry* pts,
                      signed char* dig, size_t stride, sc& t) {
    fe rx, px;
    bool ovs;
    sc s;
    bool ok = fe_set_b32(rx, sig64);           // r < p               (main_impl.h:235)
    sc_set_b32(s, sig64 + 32, &ovs);
    ok = ok && !ovs;                           // s < n               (main_impl.h:239-242)
    ok = fe_set_b32(px, xonly32) && ok;        // px < p              (extrakeys/main_impl.h:32)
    ge R, P;
    ok = ge_set_xo(R, rx, false) && ok;        // R = lift_x(r), even y
    ok = ge_set_xo(P, px, false) && ok;        // P = lift_x(px)      (extrakeys/main_impl.h:35)
    SV_UNROLL
    for (int k = 0; k < 8; k++) t.v[k] = 0;
    for (int term = 0; term < SV_SB_TERMS; term++)
        for (int w = 0; w < SV_SB_WINDOWS; w++) dig[(size_t)w * stride + term] = 0;
    if (!ok) return false;
    fe beta, bx;
    SV_UNROLL
    for (int k = 0; k < 8; k++) beta.v[k] = GE_BETA[k];
    fe_normalize(R.x); fe_normalize(R.y); fe_normalize(P.x); fe_normalize(P.y);
    fe_mul(bx, R.x, beta);
    fe_normalize(bx);
    fe_to_words(pts[0].x, R.x); fe_to_words(pts[0].y, R.y); fe_to_words(pts[0].h, bx);
    fe_mul(bx, P.x, beta);
    fe_normalize(bx);
    fe_to_words(pts[1].x, P.x); fe_to_words(pts[1].y, P.y); fe_to_words(pts[1].h, bx);
    // a = alpha + beta*lambda, (alpha, beta) = first 16 bytes of SHA-256(seed || LE64(index)), alpha forced odd
    u8 buf[40];
    for (int k = 0; k < 32; k++) buf[k] = seed32[k];
    for (int k = 0; k < 8; k++) buf[32 + k] = (u8)(index >> (8 * k));
    u32 st[8];
    sha256_bytes(st, buf, 40);
    u32 al[5] = {0, 0, 0, 0, 0}, be[5] = {0, 0, 0, 0, 0};
    al[0] = (st[0] | 1u); al[1] = st[1];
    be[0] = st[2]; be[1] = st[3];
    sc a, lam, bsc, e, c;
    SV_UNROLL
    for (int k = 0; k < 8; k++) { a.v[k] = (k < 2) ? al[k] : 0u; bsc.v[k] = (k < 2) ? be[k] : 0u; }
    sc mlam;
    SV_UNROLL
    for (int k = 0; k < 8; k++) mlam.v[k] = SC_MINUS_LAMBDA[k];
    sc_negate(lam, mlam);
    sc_mul(bsc, bsc, lam);
    sc_add(a, a, bsc);
    u8 e32[32];
    sha256_bip340_challenge(e32, sig64, xonly32, msg32);
    sc_set_b32(e, e32, nullptr);
    sc_mul(c, a, e);
    sc_mul(t, a, s);
    sb_recode(dig + 0, stride, al, 0);
    sb_recode(dig + 1, stride, be, 0);
    // GLV halves of c as sign + magnitude (|half| < 2^128, scalar_impl.h:180-282)
    sc r1, r2, tneg;
    sc_split_lambda(r1, r2, c);
    u32 m1[5], m2[5];
    u32 n1 = sc_is_high(r1) ? 1u : 0u, n2 = sc_is_high(r2) ? 1u : 0u;
    if (n1) sc_negate(tneg, r1); else tneg = r1;
    SV_UNROLL
    for (int k = 0; k < 5; k++) m1[k] = tneg.v[k];
    if (n2) sc_negate(tneg, r2); else tneg = r2;
    SV_UNROLL
    for (int k = 0; k < 5; k++) m2[k] = tneg.v[k];
    sb_recode(dig + 2, stride, m1, n1);
    sb_recode(dig + 3, stride, m2, n2);
    return true;
}

// entry e of a group = 4*i + term: point (term < 2 ? R_i : P_i), lambda form for odd terms; negative digit -> -point
SV_HD void sb_fetch(ge& p, const qtab_entry* group_pts, u32 entry, bool neg) {
    const qtab_entry* q = group_pts + 2 * (entry >> 2) + ((entry >> 1) & 1u);
    fe_from_words(p.x, (entry & 1u) ? q->h : q->x);
    fe_from_words(p.y, q->y);
    if (neg) fe_neg(p.y, p.y);
}

// window sum of one (group, window) the straightforward way (host build / cross-check): sum_b (b+1)*B_b by running sums
SV_HD void sb_window_sum_reference(gej& S, const qtab_entry* group_pts, const signed char* row, u32 entries) {
    gej B[32];
    for (int b = 0; b < 32; b++) { B[b].inf = 1; fe_set_zero(B[b].x); fe_set_zero(B[b].y); fe_set_zero(B[b].z); }
    for (u32 e = 0; e < entries; e++) {
        int d = row[e];
        if (d == 0) continue;
        ge p;
        sb_fetch(p, group_pts, e, d < 0);
        int b = (d < 0 ? -d : d) - 1;
        gej_add_ge(B[b], B[b], p);
    }
    gej run, tot;
    run.inf = 1; fe_set_zero(run.x); fe_set_zero(run.y); fe_set_zero(run.z);
    tot = run;
    for (int b = 31; b >= 0; b--) {
        gej_add_gej(run, run, B[b]);
        gej_add_gej(tot, tot, run);
    }
    S = tot;
}

// combine the window sums of one group, subtract (sum t_i)*G, test for infinity
SV_HD bool sb_group_check(const sv_jac* S, const sc* t, u32 count, const ge_mem* gtab) {
    gej T;
    small_jac_load(T, &S[SV_SB_WINDOWS - 1]);
    for (int w = SV_SB_WINDOWS - 2; w >= 0; w--) {
        for (int k = 0; k < 6; k++)
            if (!T.inf) gej_double(T, T);
        gej W;
        small_jac_load(W, &S[w]);
        gej_add_gej(T, T, W);
    }
    sc tsum, tneg;
    SV_UNROLL
    for (int k = 0; k < 8; k++) tsum.v[k] = 0;
    for (u32 i = 0; i < count; i++) sc_add(tsum, tsum, t[i]);
    sc_negate(tneg, tsum);
    if (!sc_is_zero(tneg)) {
        sv_work w;
        sc_prepare_u1(w, tneg);
        for (int row = 0; row < 16; row++) {
            int d = w.gd[row];
            if (d != 0) {
                ge p;
                u32 a = (u32)(d < 0 ? -d : d);
                ge_from_mem(p, gtab + (size_t)row * SV_GT_ROW + (a - 1));
                if (d < 0) fe_neg(p.y, p.y);
                gej_add_ge(T, T, p);
            }
        }
    }
    return T.inf != 0;
}
