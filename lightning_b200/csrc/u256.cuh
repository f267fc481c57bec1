// u256.cuh — 256-bit integer primitives on 8x32-bit little-endian limbs.
//
// Device path: inline-PTX carry chains.  `mad.lo.cc.u32` + `madc.hi.cc.u32` on the same operand
// pair are fused by ptxas into ONE `IMAD.WIDE.U32(.X)` with the carry in a predicate register
// (checked with cuobjdump -sass for sm_100a: a full 8x8-limb product is 64 IMAD.WIDE + ~23
// IADD3/SEL).  The product is accumulated in two interleaved 64-bit-slot arrays (even/odd
// column alignment) so that every IMAD.WIDE lands on an aligned (lo,hi) pair and carry chains
// run 4 deep per row; the two arrays are merged with a single 15-limb add chain.
//
// Host path (tests/host_emul only): plain uint64_t arithmetic with identical semantics.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// add / sub with carry, 8 limbs
// ---------------------------------------------------------------------------------------------
SV_HD u32 u256_add(u32 r[8], const u32 a[8], const u32 b[8]) {
#if SV_DEVICE_CODE
    u32 c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    u64 c = 0;
    for (int i = 0; i < 8; i++) { c += (u64)a[i] + b[i]; r[i] = (u32)c; c >>= 32; }
    return (u32)c;
#endif
}

SV_HD u32 u256_sub(u32 r[8], const u32 a[8], const u32 b[8]) {
#if SV_DEVICE_CODE
    u32 bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return bw & 1u;  // subc of 0-0-borrow gives 0xFFFFFFFF when borrow
#else
    u64 bw = 0;
    for (int i = 0; i < 8; i++) {
        u64 d = (u64)a[i] - b[i] - bw;
        r[i] = (u32)d;
        bw = (d >> 32) & 1;
    }
    return (u32)bw;
#endif
}

// compare: a >= b ?
SV_HD bool u256_gte(const u32 a[8], const u32 b[8]) {
    u32 t[8];
    return u256_sub(t, a, b) == 0;
}
SV_HD bool u256_is_zero(const u32 a[8]) {
    return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0;
}
SV_HD bool u256_eq(const u32 a[8], const u32 b[8]) {
    u32 d = 0;
    SV_UNROLL
    for (int i = 0; i < 8; i++) d |= a[i] ^ b[i];
    return d == 0;
}

// ---------------------------------------------------------------------------------------------
// 256x256 -> 512 product
// ---------------------------------------------------------------------------------------------
#if SV_DEVICE_CODE
// acc[0..7] += {a0,a1,a2,a3} * b as four chained 64-bit multiply-accumulates; returns carry-out.
SV_D u32 sv_cmad4(u32* acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 b) {
    u32 c;
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=r"(c)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
    return c;
}
// r[0] = e[0]; r[k] = e[k] + o[k-1] (+carry), k = 1..15   (merge of the even/odd column arrays)
SV_D void sv_merge16(u32 r[16], const u32 e[16], const u32 o[16]) {
    r[0] = e[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
          "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
          "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]));
}
#endif

SV_HD void u256_mul_wide(u32 r[16], const u32 a[8], const u32 b[8]) {
#if SV_DEVICE_CODE
    // E holds products whose low limb sits at an even position, O those at an odd position
    // (O[k] is limb position k+1).  Zero-initialised; ptxas folds the zeros into RZ operands.
    u32 E[18], O[18];
    SV_UNROLL
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
    SV_UNROLL
    for (int i = 0; i < 8; i++) {
        u32* A = (i & 1) ? (O + i - 1) : (E + i);  // a[even j] * b[i] -> position i+j (parity of i)
        u32* B = (i & 1) ? (E + i + 1) : (O + i);  // a[odd j]  * b[i] -> position i+j (parity of i+1)
        u32 c = sv_cmad4(A, a[0], a[2], a[4], a[6], b[i]);
        A[8] = c;  // limb i+8 of that array is still untouched at this point
        (void)sv_cmad4(B, a[1], a[3], a[5], a[7], b[i]);  // top product lands on fresh limbs: no carry-out
    }
    sv_merge16(r, E, O);
#else
    u64 t[16];
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        u64 c = 0;
        for (int j = 0; j < 8; j++) {
            u64 p = (u64)a[j] * b[i] + t[i + j] + c;
            t[i + j] = (u32)p;
            c = p >> 32;
        }
        t[i + 8] = c;
    }
    for (int i = 0; i < 16; i++) r[i] = (u32)t[i];
#endif
}

SV_HD void u256_sqr_wide(u32 r[16], const u32 a[8]) {
    // TODO(perf): dedicated squaring (36 products instead of 64)
    u256_mul_wide(r, a, a);
}
