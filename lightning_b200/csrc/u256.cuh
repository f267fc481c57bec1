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
#include "u256_gen.cuh"  // generated PTX bodies: sv_mul8_dev, sv_sqr8_dev (tools/gen_mul.py)

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

// Schoolbook 8x8 (64 IMAD.WIDE).  Kept as the reference implementation of the product and used by
// the scalar-field code; the field code uses the Karatsuba / dedicated-square versions below.
SV_HD void u256_mul_wide_schoolbook(u32 r[16], const u32 a[8], const u32 b[8]) {
#if SV_DEVICE_CODE
    sv_mul8_dev(r, a, b);
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

// ---------------------------------------------------------------------------------------------
// 128x128 -> 256 product (16 IMAD.WIDE), building block of the Karatsuba multiply
// ---------------------------------------------------------------------------------------------
#if SV_DEVICE_CODE
SV_D u32 sv_cmad2(u32* acc, u32 a0, u32 a1, u32 b) {  // acc[0..3] += {a0,a1}*b ; returns carry-out
    u32 c;
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(c)
        : "r"(a0), "r"(a1), "r"(b));
    return c;
}
#endif
SV_HD void u128_mul_wide(u32 r[8], const u32 a[4], const u32 b[4]) {
#if SV_DEVICE_CODE
    u32 E[10], O[10];
    SV_UNROLL
    for (int i = 0; i < 10; i++) { E[i] = 0; O[i] = 0; }
    SV_UNROLL
    for (int i = 0; i < 4; i++) {
        u32* A = (i & 1) ? (O + i - 1) : (E + i);
        u32* B = (i & 1) ? (E + i + 1) : (O + i);
        u32 c = sv_cmad2(A, a[0], a[2], b[i]);
        A[4] = c;
        (void)sv_cmad2(B, a[1], a[3], b[i]);
    }
    r[0] = E[0];
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
#else
    u64 t[8];
    for (int i = 0; i < 8; i++) t[i] = 0;
    for (int i = 0; i < 4; i++) {
        u64 c = 0;
        for (int j = 0; j < 4; j++) {
            u64 p = (u64)a[j] * b[i] + t[i + j] + c;
            t[i + j] = (u32)p;
            c = p >> 32;
        }
        t[i + 4] = c;
    }
    for (int i = 0; i < 8; i++) r[i] = (u32)t[i];
#endif
}

// |a - b| for 128-bit operands; returns 1 if a < b (i.e. the result was negated)
SV_HD u32 u128_absdiff(u32 r[4], const u32 a[4], const u32 b[4]) {
#if SV_DEVICE_CODE
    u32 bw;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    // bw = 0 or 0xFFFFFFFF ; conditional negate: (x ^ bw) - bw
    u32 one = bw & 1u;
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, 0;\n\t"
        "addc.cc.u32 %2, %6, 0;\n\t"
        "addc.u32 %3, %7, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
        : "r"(r[0] ^ bw), "r"(r[1] ^ bw), "r"(r[2] ^ bw), "r"(r[3] ^ bw), "r"(one));
    return one;
#else
    u64 bw = 0;
    u32 t[4];
    for (int i = 0; i < 4; i++) {
        u64 d = (u64)a[i] - b[i] - bw;
        t[i] = (u32)d;
        bw = (d >> 32) & 1;
    }
    u32 mask = bw ? 0xFFFFFFFFu : 0u;
    u64 c = bw;
    for (int i = 0; i < 4; i++) {
        c += (u64)(t[i] ^ mask);
        r[i] = (u32)c;
        c >>= 32;
    }
    return (u32)bw;
#endif
}

// Karatsuba, one level, subtractive form:
//   a = a0 + a1 W, b = b0 + b1 W (W = 2^128):  ab = z0 + (z0 + z2 + s |a0-a1| |b1-b0|) W + z2 W^2
// 48 IMAD.WIDE instead of 64.  On B200 the 64-bit IMAD.WIDE issues at half the rate of IADD3 (measured,
// see profiles/), so trading 16 multiplies for ~60 carry-chain adds shortens the critical pipe.
SV_HD void u256_mul_wide_karatsuba(u32 r[16], const u32 a[8], const u32 b[8]) {
    u32 z0[8], z2[8], m[8], da[4], db[4];
    u128_mul_wide(z0, a, b);
    u128_mul_wide(z2, a + 4, b + 4);
    u32 sa = u128_absdiff(da, a, a + 4);      // a0 - a1
    u32 sb = u128_absdiff(db, b + 4, b);      // b1 - b0
    u128_mul_wide(m, da, db);
    u32 neg = sa ^ sb;                        // middle term is z0 + z2 - m when the signs differ
#if SV_DEVICE_CODE
    u32 mask = 0u - neg;
    u32 t[9], z1[9];
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8])
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]),
          "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
    // z1 = t + (m ^ mask) + neg, ninth limb absorbs the two's-complement wrap (+mask)
    asm("add.cc.u32 %0, %18, 0xFFFFFFFF;\n\t"   // carry := neg
        "addc.cc.u32 %0, %9, %19;\n\t"
        "addc.cc.u32 %1, %10, %20;\n\t"
        "addc.cc.u32 %2, %11, %21;\n\t"
        "addc.cc.u32 %3, %12, %22;\n\t"
        "addc.cc.u32 %4, %13, %23;\n\t"
        "addc.cc.u32 %5, %14, %24;\n\t"
        "addc.cc.u32 %6, %15, %25;\n\t"
        "addc.cc.u32 %7, %16, %26;\n\t"
        "addc.u32 %8, %17, %27;"
        : "=&r"(z1[0]), "=&r"(z1[1]), "=&r"(z1[2]), "=&r"(z1[3]), "=&r"(z1[4]), "=&r"(z1[5]), "=&r"(z1[6]), "=&r"(z1[7]),
          "=&r"(z1[8])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]),
          "r"(neg), "r"(m[0] ^ mask), "r"(m[1] ^ mask), "r"(m[2] ^ mask), "r"(m[3] ^ mask), "r"(m[4] ^ mask),
          "r"(m[5] ^ mask), "r"(m[6] ^ mask), "r"(m[7] ^ mask), "r"(mask));
    // r = z0 + z1 W + z2 W^2
    r[0] = z0[0]; r[1] = z0[1]; r[2] = z0[2]; r[3] = z0[3];
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, 0;\n\t"
        "addc.cc.u32 %10, %22, 0;\n\t"
        "addc.u32 %11, %23, 0;"
        : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]),
          "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]),
          "r"(z1[0]), "r"(z1[1]), "r"(z1[2]), "r"(z1[3]), "r"(z1[4]), "r"(z1[5]), "r"(z1[6]), "r"(z1[7]), "r"(z1[8]));
#else
    u32 mask = neg ? 0xFFFFFFFFu : 0u;
    u32 t[9], z1[9];
    u64 c = 0;
    for (int i = 0; i < 8; i++) { c += (u64)z0[i] + z2[i]; t[i] = (u32)c; c >>= 32; }
    t[8] = (u32)c;
    c = neg;
    for (int i = 0; i < 8; i++) { c += (u64)t[i] + (m[i] ^ mask); z1[i] = (u32)c; c >>= 32; }
    z1[8] = (u32)((u64)t[8] + mask + c);
    for (int i = 0; i < 4; i++) r[i] = z0[i];
    c = 0;
    for (int i = 0; i < 12; i++) {
        u32 base = (i < 4) ? z0[4 + i] : z2[i - 4];
        c += (u64)base + (i < 9 ? z1[i] : 0u);
        r[4 + i] = (u32)c;
        c >>= 32;
    }
#endif
}

// Measured on B200 (profiles/r1_probes.md): IMAD.WIDE.U32 issues at 32 lanes/clk/SM (one warp instruction per
// 4 cycles per SM sub-partition), IADD3 at 64 lanes/clk/SM.  Schoolbook: 73 IMAD.WIDE + ~50 other -> bound by
// the multiplier pipe at ~292 cycles; Karatsuba: 56 IMAD.WIDE + ~170 other -> ~282 cycles when ptxas balances
// the two pipes, and measured SLOWER (9.8e10 vs 1.09e11 field mults/s).  So the product stays schoolbook and
// Karatsuba is kept only as a tested alternative.
SV_HD void u256_mul_wide(u32 r[16], const u32 a[8], const u32 b[8]) { u256_mul_wide_schoolbook(r, a, b); }

// Dedicated squaring: 28 cross products (doubled) + 8 diagonal squares = 36 IMAD.WIDE instead of 64.
SV_HD void u256_sqr_wide(u32 r[16], const u32 a[8]) {
#if SV_DEVICE_CODE
    sv_sqr8_dev(r, a);
#else
    u256_mul_wide_schoolbook(r, a, a);
#endif
}

// ---------------------------------------------------------------------------------------------
// a^-1 mod m for an odd modulus m (a < m; 0 -> 0), variable time: binary extended Euclid with the invariants
// x1*a == u and x2*a == v (mod m).  Every pass halves u (after making it even by subtracting the smaller of u, v), so the
// loop ends within bits(u) + bits(v) <= 512 passes.  Used where ONE inversion sits on the critical path of a lone
// verification (small-batch path): ~500 short passes instead of the ~330 dependent multiplications of a Fermat chain.
// The reference inverts with safegcd (modinv64_impl.h:638), also variable time in the verification path.
// ---------------------------------------------------------------------------------------------
SV_HD void u256_shr1(u32 r[8], u32 top) {  // r = (top:r) >> 1
    SV_UNROLL
    for (int i = 0; i < 7; i++) r[i] = (r[i] >> 1) | (r[i + 1] << 31);
    r[7] = (r[7] >> 1) | (top << 31);
}
SV_HD void u256_modinv_var(u32 r[8], const u32 a[8], const u32 m[8]) {
    u32 u[8], v[8], x1[8], x2[8];
    SV_UNROLL
    for (int i = 0; i < 8; i++) { u[i] = a[i]; v[i] = m[i]; x1[i] = (i == 0); x2[i] = 0; }
#if SV_DEVICE_CODE
#pragma unroll 1
#endif
    for (int it = 0; it < 520; it++) {
        if (u256_is_zero(u)) break;
        if (u[0] & 1u) {
            u32 t[8];
            u32 lt = u256_sub(t, u, v);  // borrow <=> u < v
            if (lt) {                    // swap the pairs, then u - v is v_old - u_old = -t
                SV_UNROLL
                for (int i = 0; i < 8; i++) { u32 w = u[i]; u[i] = v[i]; v[i] = w; w = x1[i]; x1[i] = x2[i]; x2[i] = w; }
                (void)u256_sub(u, u, v);
            } else {
                SV_UNROLL
                for (int i = 0; i < 8; i++) u[i] = t[i];
            }
            u32 bw = u256_sub(x1, x1, x2);  // x1 = x1 - x2 mod m
            if (bw) (void)u256_add(x1, x1, m);
        }
        u256_shr1(u, 0);
        u32 c = 0;
        if (x1[0] & 1u) c = u256_add(x1, x1, m);  // make x1 even (m is odd), keeping the 257th bit
        u256_shr1(x1, c);
    }
    SV_UNROLL
    for (int i = 0; i < 8; i++) r[i] = x2[i];
}
