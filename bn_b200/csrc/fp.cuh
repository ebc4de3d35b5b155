// fp.cuh -- 256-bit Montgomery prime-field arithmetic for sm_100a, 8 x 32-bit limbs in registers.
//
// Replaces (bit-exactly, on canonical Montgomery values) the reference's src/arith.rs:238-273, 481-503
// (U256 add/sub/neg/mul + mul_reduce) and src/fields/fp.rs:15-22, 103-157 (Fq / Fr newtypes).
//
// Design (B200-first, not a translation of the reference's 4x64 word-serial loop):
//   * B200 has no 64-bit integer multiplier; the unit of work is IMAD.WIDE.U32 (32x32+64 -> 64) on the
//     fma pipe.  A mad.lo.cc/madc.hi.cc pair with a pair-aligned accumulator fuses into ONE
//     IMAD.WIDE.U32.X with carry-in/out, so products are accumulated in two interleaved limb arrays
//     ("even" E and "odd" O, value = E + 2^32*O) whose register pairs never straddle an IMAD.WIDE result.
//   * Multiplication is "separated": full 512-bit product (64 IMAD.WIDE) then a word-serial Montgomery
//     reduction (64 IMAD.WIDE + 8 IMAD).  Keeping the two apart is what lets the tower code add several
//     512-bit products BEFORE one reduction (lazy reduction): the pairing kernel spends 2 reductions per
//     Fq2 output coefficient instead of one per Fq product.
//   * Every value that leaves a function named fp_* is canonical in [0, p).
//
// The carry chains are single inline-PTX statements (one statement == one chain, so the CC flag never
// crosses a statement boundary).  A portable C fallback of the same primitives exists for host builds:
// tests/host_emu compiles the kernels' tower / pairing logic with g++ and checks it against the oracle
// without a GPU.  The product library itself never runs that path (see kernels.cu: no CPU fallback).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BN_HD __host__ __device__ __forceinline__
#define BN_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BN_HD inline
#define BN_HD_NOINLINE inline
#endif

#ifndef BN_MUL_FUSED
#define BN_MUL_FUSED 1   // fp_mul / fp_mul2: product and reduction rows interleaved on one accumulator (0: product, merge, reduce)
#endif

#if defined(__CUDA_ARCH__)
#define BN_UNROLL _Pragma("unroll")
#else
#define BN_UNROLL
#endif

namespace bn {

// ------------------------------------------------------------------------------------------------
// Moduli.  Values derived in gen_constants.py (q = 36u^4+36u^3+24u^2+6u+1, r = ...+18u^2...);
// INV = -p^-1 mod 2^32.
// ------------------------------------------------------------------------------------------------
#include "constants_fp.inc"

// ------------------------------------------------------------------------------------------------
// Carry-chain primitives
// ------------------------------------------------------------------------------------------------
// BN_MAD_VOLATILE=1 pins the multiply chains in program order (asm volatile): one dependent IMAD.WIDE.X chain already
// issues at the fma-heavy pipe's rate (4 cycles per warp instruction), so interleaving chains buys nothing and costs
// ptxas predicate spills (LOP3 into a mask register) and register moves.
#ifndef BN_MAD_VOLATILE
#define BN_MAD_VOLATILE 0
#endif
#if BN_MAD_VOLATILE
#define BN_MAD_ASM asm volatile
#else
#define BN_MAD_ASM asm
#endif

// acc[0..7] += {x0,x1,x2,x3} * y laid out as four (lo,hi) pairs; carry-out added into acc[8].
BN_HD void mad_row4(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if defined(__CUDA_ARCH__)
    BN_MAD_ASM("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(acc[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
        uint64_t p = (uint64_t)x[j] * y;
        uint64_t t = (uint64_t)acc[2 * j] + (uint32_t)p + c;
        acc[2 * j] = (uint32_t)t;
        t = (uint64_t)acc[2 * j + 1] + (p >> 32) + (t >> 32);
        acc[2 * j + 1] = (uint32_t)t;
        c = t >> 32;
    }
    acc[8] += (uint32_t)c;
#endif
}

// Shorter chains for the squaring triangle: acc[0 .. 2N) += {x0 .. x(N-1)} * y, carry-out added into acc[2N].
BN_HD void mad_row3(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
#if defined(__CUDA_ARCH__)
    BN_MAD_ASM("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
#else
    const uint32_t x[3] = {x0, x1, x2};
    uint64_t c = 0;
    for (int j = 0; j < 3; j++) {
        uint64_t p = (uint64_t)x[j] * y;
        uint64_t t = (uint64_t)acc[2 * j] + (uint32_t)p + c;
        acc[2 * j] = (uint32_t)t;
        t = (uint64_t)acc[2 * j + 1] + (p >> 32) + (t >> 32);
        acc[2 * j + 1] = (uint32_t)t;
        c = t >> 32;
    }
    acc[6] += (uint32_t)c;
#endif
}
BN_HD void mad_row2(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t y) {
#if defined(__CUDA_ARCH__)
    BN_MAD_ASM("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
        : "r"(x0), "r"(x1), "r"(y));
#else
    const uint32_t x[2] = {x0, x1};
    uint64_t c = 0;
    for (int j = 0; j < 2; j++) {
        uint64_t p = (uint64_t)x[j] * y;
        uint64_t t = (uint64_t)acc[2 * j] + (uint32_t)p + c;
        acc[2 * j] = (uint32_t)t;
        t = (uint64_t)acc[2 * j + 1] + (p >> 32) + (t >> 32);
        acc[2 * j + 1] = (uint32_t)t;
        c = t >> 32;
    }
    acc[4] += (uint32_t)c;
#endif
}
BN_HD void mad_row1(uint32_t* acc, uint32_t x0, uint32_t y) {
#if defined(__CUDA_ARCH__)
    BN_MAD_ASM("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2])
        : "r"(x0), "r"(y));
#else
    uint64_t p = (uint64_t)x0 * y;
    uint64_t t = (uint64_t)acc[0] + (uint32_t)p;
    acc[0] = (uint32_t)t;
    t = (uint64_t)acc[1] + (p >> 32) + (t >> 32);
    acc[1] = (uint32_t)t;
    acc[2] += (uint32_t)(t >> 32);
#endif
}

// Same, without a carry-out limb (caller proved the carry is zero).
BN_HD void mad_row4_nc(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if defined(__CUDA_ARCH__)
    BN_MAD_ASM("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    uint32_t t[9];
    for (int j = 0; j < 8; j++) t[j] = acc[j];
    t[8] = 0;
    mad_row4(t, x0, x1, x2, x3, y);
    for (int j = 0; j < 8; j++) acc[j] = t[j];
#endif
}

// Carry-in = carry of (f0 + f1) (the two halves of an already-cancelled limb), then as mad_row4.
BN_HD void mad_row4_fold(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y,
                         uint32_t f0, uint32_t f1, bool carry_out) {
#if defined(__CUDA_ARCH__)
    uint32_t scratch = 0;
    if (carry_out) {
        asm("add.cc.u32 %9, %15, %16;\n\t"
            "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
            "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
            "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
            "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
            "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
            "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
            "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
            "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
              "+r"(acc[7]), "+r"(acc[8]), "=&r"(scratch)
            : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y), "r"(f0), "r"(f1));
    } else {
        asm("add.cc.u32 %8, %14, %15;\n\t"
            "madc.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.u32 %7, %12, %13, %7;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
              "+r"(acc[7]), "=&r"(scratch)
            : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y), "r"(f0), "r"(f1));
    }
    (void)scratch;
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t c = ((uint64_t)f0 + f1) >> 32;
    for (int j = 0; j < 4; j++) {
        uint64_t p = (uint64_t)x[j] * y;
        uint64_t t = (uint64_t)acc[2 * j] + (uint32_t)p + c;
        acc[2 * j] = (uint32_t)t;
        t = (uint64_t)acc[2 * j + 1] + (p >> 32) + (t >> 32);
        acc[2 * j + 1] = (uint32_t)t;
        c = t >> 32;
    }
    if (carry_out) acc[8] += (uint32_t)c;
#endif
}

// r[0..7] = a[0..7] + b[0..7]; returns carry-out.
BN_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}

// r[0..7] = a[0..7] + b[0..7] + cin (cin in {0,1}); returns carry-out.
BN_HD uint32_t add8c(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t cin) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"  // CF = cin
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
#else
    uint64_t t = cin;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}

// r[0..7] = a[0..7] - b[0..7]; returns borrow (1 if a < b).
BN_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t bw;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    bw &= 1u;  // subc 0-0-borrow = 0xffffffff when borrow
#else
    int64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)t;
        t >>= 32;  // arithmetic: -1 on borrow
    }
    bw = (uint32_t)(t & 1);
#endif
    return bw;
}

// acc[0..7] += b[0..7] (+cin); returns carry-out.  In-place form: accumulators are tied "+r" operands.
BN_HD uint32_t addi8(uint32_t* acc, const uint32_t* b) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=&r"(c)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)acc[i] + b[i];
        acc[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}
BN_HD uint32_t addi8c(uint32_t* acc, const uint32_t* b, uint32_t cin) {
    uint32_t c = cin;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %8, %8, 0xffffffff;\n\t"  // CF = cin (cin in {0,1})
        "addc.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(c)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t t = cin;
    for (int i = 0; i < 8; i++) {
        t += (uint64_t)acc[i] + b[i];
        acc[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}

// acc[0..7] -= b[0..7] (- bin); returns borrow-out (0/1).  In place.
BN_HD uint32_t subi8b(uint32_t* acc, const uint32_t* b, uint32_t bin) {
    uint32_t bo = bin;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %8, 0, %8;\n\t"  // CF(borrow) = (bin != 0); bin in {0,1}
        "subc.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(bo)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    bo &= 1u;
#else
    int64_t t = -(int64_t)bin;
    for (int i = 0; i < 8; i++) {
        t += (int64_t)acc[i] - (int64_t)b[i];
        acc[i] = (uint32_t)t;
        t >>= 32;
    }
    bo = (uint32_t)(t & 1);
#endif
    return bo;
}
// acc[0..15] -= b[0..15] (mod 2^512)
BN_HD void sub16(uint32_t* acc, const uint32_t* b) {
    uint32_t bo = subi8b(acc, b, 0u);
    (void)subi8b(acc + 8, b + 8, bo);
}

// 16-limb accumulate: acc[0..15] += b[0..15] (carry-out dropped: callers keep totals < 2^512).
BN_HD void add16(uint32_t* acc, const uint32_t* b) {
    uint32_t c = addi8(acc, b);
    (void)addi8c(acc + 8, b + 8, c);
}
// acc[1..15] += b[0..14]  (i.e. acc += b << 32, b[15] known to be zero)
BN_HD void add16_shift1(uint32_t* acc, const uint32_t* b) {
    uint32_t c = addi8(acc + 1, b);
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %7, %7, 0xffffffff;\n\t"  // CF = c (c in {0,1}); %7 is scratch copy of c
        "addc.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.u32 %6, %6, %14;"
        : "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(acc[12]), "+r"(acc[13]), "+r"(acc[14]), "+r"(acc[15]),
          "+r"(c)
        : "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]));
#else
    uint64_t t = c;
    for (int i = 0; i < 7; i++) {
        t += (uint64_t)acc[9 + i] + b[8 + i];
        acc[9 + i] = (uint32_t)t;
        t >>= 32;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// Field element + modulus descriptors
// ------------------------------------------------------------------------------------------------
struct alignas(16) Fp {
    uint32_t v[8];
};

struct ModQ {  // base field Fq   (reference src/fields/fp.rs:170-177)
    static BN_HD uint32_t m(int i) { return FQ_MOD_f(i); }
    static BN_HD uint32_t m2(int i) { return FQ_MOD2_f(i); }  // 2q
    static BN_HD uint32_t inv() { return FQ_INV32; }
};
struct ModR {  // scalar field Fr (reference src/fields/fp.rs:161-168)
    static BN_HD uint32_t m(int i) { return FR_MOD_f(i); }
    static BN_HD uint32_t m2(int i) { return FR_MOD2_f(i); }
    static BN_HD uint32_t inv() { return FR_INV32; }
};

template <class M>
BN_HD void load_mod(uint32_t* p) {
    BN_UNROLL
    for (int i = 0; i < 8; i++) p[i] = M::m(i);
}
template <class M>
BN_HD void load_mod2(uint32_t* p) {
    BN_UNROLL
    for (int i = 0; i < 8; i++) p[i] = M::m2(i);
}

BN_HD bool fp_is_zero(const Fp& a) {
    uint32_t o = 0;
    BN_UNROLL
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
BN_HD bool fp_eq(const Fp& a, const Fp& b) {
    uint32_t o = 0;
    BN_UNROLL
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
BN_HD Fp fp_zero() {
    Fp r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
BN_HD Fp fp_select(bool c, const Fp& a, const Fp& b) {  // c ? a : b
    Fp r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

// if (x >= k*p) x -= k*p, for k*p given as limbs
BN_HD void cond_sub_limbs(uint32_t* x, const uint32_t* kp) {
    uint32_t t[8];
    uint32_t bw = sub8(t, x, kp);
    BN_UNROLL
    for (int i = 0; i < 8; i++) x[i] = bw ? x[i] : t[i];
}
template <class M>
BN_HD void cond_sub_p(uint32_t* x) {
    uint32_t p[8];
    load_mod<M>(p);
    cond_sub_limbs(x, p);
}
template <class M>
BN_HD void cond_sub_2p(uint32_t* x) {
    uint32_t p[8];
    load_mod2<M>(p);
    cond_sub_limbs(x, p);
}

// (a + b) mod p, inputs canonical.   reference src/arith.rs:238-244
template <class M>
BN_HD Fp fp_add(const Fp& a, const Fp& b) {
    Fp r;
    (void)add8(r.v, a.v, b.v);  // < 2p < 2^255: no carry
    cond_sub_p<M>(r.v);
    return r;
}
// (a - b) mod p.   reference src/arith.rs:247-253
template <class M>
BN_HD Fp fp_sub(const Fp& a, const Fp& b) {
    Fp r;
    uint32_t t[8], p[8];
    uint32_t bw = sub8(r.v, a.v, b.v);
    load_mod<M>(p);
    (void)add8(t, r.v, p);
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = bw ? t[i] : r.v[i];
    return r;
}
// -a mod p (0 stays 0).   reference src/arith.rs:266-273
template <class M>
BN_HD Fp fp_neg(const Fp& a) {
    Fp r;
    uint32_t p[8];
    load_mod<M>(p);
    (void)sub8(r.v, p, a.v);
    bool z = fp_is_zero(a);
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
    return r;
}
// p - a WITHOUT the zero fix-up: result in [1, p]; only valid as a multiplier / addend feeding a reduction.
template <class M>
BN_HD Fp fp_neg_lazy(const Fp& a) {
    Fp r;
    uint32_t p[8];
    load_mod<M>(p);
    (void)sub8(r.v, p, a.v);
    return r;
}
// The modulus held in REGISTERS, loaded from memory so that ptxas cannot fold it back into immediates: p - a then is one
// IADD3.X per limb (register operands take the ~ modifier); with immediate modulus limbs ptxas spends a LOP3 + IADD3.X
// per limb.  Worth it when one function negates several values (hx_cyc_sqr: six).
struct ModRegs {
    uint32_t p[8];
};
template <class M>
BN_HD ModRegs mod_regs() {  // portable form (immediates); kernels load the limbs from shared memory instead (Ctx::mod_q)
    ModRegs r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.p[i] = M::m(i);
    return r;
}
BN_HD Fp fp_neg_lazy_r(const ModRegs& q, const Fp& a) {
    Fp r;
    (void)sub8(r.v, q.p, a.v);
    return r;
}
template <class M>
BN_HD Fp fp_dbl(const Fp& a) {
    return fp_add<M>(a, a);
}
// a/2 mod p
template <class M>
BN_HD Fp fp_half(const Fp& a) {
    uint32_t t[8], p[8];
    load_mod<M>(p);
    bool odd = a.v[0] & 1u;
    BN_UNROLL
    for (int i = 0; i < 8; i++) p[i] = odd ? p[i] : 0u;
    (void)add8(t, a.v, p);  // < 2p < 2^255
    Fp r;
    BN_UNROLL
    for (int i = 0; i < 7; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r.v[7] = t[7] >> 1;
    return r;
}

// ------------------------------------------------------------------------------------------------
// 512-bit products in (E, O) form and their accumulation
// ------------------------------------------------------------------------------------------------
struct alignas(16) Wide {
    uint32_t w[16];
};
BN_HD Wide wide_zero() {
    Wide r;
    BN_UNROLL
    for (int i = 0; i < 16; i++) r.w[i] = 0;
    return r;
}

// One row (multiplier limb y = index i) of a*y into the (E,O) pair.  E has 17 slots, O 16 (tops stay 0).
BN_HD void eo_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t y, int i) {
    if ((i & 1) == 0) {
        mad_row4(&E[i], a[0], a[2], a[4], a[6], y);
        mad_row4(&O[i], a[1], a[3], a[5], a[7], y);
    } else {
        mad_row4(&O[i - 1], a[0], a[2], a[4], a[6], y);
        mad_row4(&E[i + 1], a[1], a[3], a[5], a[7], y);
    }
}

// acc += a*b  (one product).  64 IMAD.WIDE.
BN_HD void wide_mac1(Wide& acc, const Fp& a, const Fp& b) {
    uint32_t E[18], O[16];
    BN_UNROLL
    for (int i = 0; i < 18; i++) E[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 16; i++) O[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 8; i++) eo_row(E, O, a.v, b.v[i], i);
    add16(acc.w, E);
    add16_shift1(acc.w, O);
}
// acc += a*b + c*d  (row-major so every carry lands in a not-yet-multiplied limb).  128 IMAD.WIDE.
BN_HD void wide_mac2(Wide& acc, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    uint32_t E[18], O[16];
    BN_UNROLL
    for (int i = 0; i < 18; i++) E[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 16; i++) O[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        eo_row(E, O, a.v, b.v[i], i);
        eo_row(E, O, c.v, d.v[i], i);
    }
    add16(acc.w, E);
    add16_shift1(acc.w, O);
}
// T = a*b as a merged 512-bit integer (fresh E/O pair, then T = E + 2^32 O).  64 IMAD.WIDE.
BN_HD void wide_mul(Wide& T, const Fp& a, const Fp& b) {
    uint32_t E[18], O[16];
    BN_UNROLL
    for (int i = 0; i < 18; i++) E[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 16; i++) O[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 8; i++) eo_row(E, O, a.v, b.v[i], i);
    BN_UNROLL
    for (int i = 0; i < 16; i++) T.w[i] = E[i];
    add16_shift1(T.w, O);
}
// T = a^2 as a 512-bit integer: the 28 cross products a_i a_j (i < j) in the (E, O) layout of eo_row, doubled, plus the
// eight squares a_i^2, which tile the 512 bits exactly (a_i^2 occupies limbs 2i, 2i + 1): 36 IMAD.WIDE against 64.
BN_HD void wide_sqr(Wide& T, const Fp& a) {
    uint32_t E[18], O[16];
    BN_UNROLL
    for (int i = 0; i < 18; i++) E[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 16; i++) O[i] = 0;
    const uint32_t* v = a.v;
    // row i multiplies by y = a_i and takes the limbs j > i; a product a_j y lands at limb position i + j: even
    // positions live in E (index = position), odd ones in O (index = position - 1)
    mad_row3(&E[2], v[2], v[4], v[6], v[0]);          // i = 0: j = 2, 4, 6   -> positions 2, 4, 6
    mad_row4(&O[0], v[1], v[3], v[5], v[7], v[0]);    //        j = 1, 3, 5, 7 -> positions 1, 3, 5, 7
    mad_row3(&O[2], v[2], v[4], v[6], v[1]);          // i = 1: j = 2, 4, 6   -> positions 3, 5, 7
    mad_row3(&E[4], v[3], v[5], v[7], v[1]);          //        j = 3, 5, 7   -> positions 4, 6, 8
    mad_row2(&E[6], v[4], v[6], v[2]);                // i = 2: j = 4, 6      -> positions 6, 8
    mad_row3(&O[4], v[3], v[5], v[7], v[2]);          //        j = 3, 5, 7   -> positions 5, 7, 9
    mad_row2(&O[6], v[4], v[6], v[3]);                // i = 3: j = 4, 6      -> positions 7, 9
    mad_row2(&E[8], v[5], v[7], v[3]);                //        j = 5, 7      -> positions 8, 10
    mad_row1(&E[10], v[6], v[4]);                     // i = 4: j = 6         -> position 10
    mad_row2(&O[8], v[5], v[7], v[4]);                //        j = 5, 7      -> positions 9, 11
    mad_row1(&O[10], v[6], v[5]);                     // i = 5: j = 6         -> position 11
    mad_row1(&E[12], v[7], v[5]);                     //        j = 7         -> position 12
    mad_row1(&O[12], v[7], v[6]);                     // i = 6: j = 7         -> position 13
    // cross = E + 2^32 O (< 2^511), doubled
    uint32_t c[16];
    BN_UNROLL
    for (int i = 0; i < 16; i++) c[i] = E[i];
    add16_shift1(c, O);
    BN_UNROLL
    for (int i = 15; i > 0; i--) c[i] = (c[i] << 1) | (c[i - 1] >> 31);
    c[0] <<= 1;
    // diagonal: eight independent 64-bit squares
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        const uint64_t d = (uint64_t)v[i] * v[i];
        T.w[2 * i] = (uint32_t)d;
        T.w[2 * i + 1] = (uint32_t)(d >> 32);
    }
    add16(T.w, c);
}
// acc <<= 1 (value doubles; caller keeps it < 2^512)
BN_HD void wide_dbl(Wide& acc) {
    BN_UNROLL
    for (int i = 15; i > 0; i--) acc.w[i] = (acc.w[i] << 1) | (acc.w[i - 1] >> 31);
    acc.w[0] <<= 1;
}

// ------------------------------------------------------------------------------------------------
// Carry-save 512-bit accumulator: products keep accumulating in the (E, O) limb arrays across MANY
// multiply-accumulates; the carry out of every 4-IMAD chain goes to a small per-position counter instead of
// rippling through limbs that already hold data.  One merge (E + 2^32 O + counters) per reduction.
// Saves the two 16-limb merges per product pair that Wide/wide_mac2 pay.
// ------------------------------------------------------------------------------------------------
struct AccEO {
    uint32_t E[16];
    uint32_t O[16];   // O[15] stays 0
    uint32_t CE[4];   // carries into E positions 8,10,12,14
    uint32_t CO[4];   // carries into O positions 8,10,12,14
};
BN_HD void acc_zero(AccEO& a) {
    BN_UNROLL
    for (int i = 0; i < 16; i++) { a.E[i] = 0; a.O[i] = 0; }
    BN_UNROLL
    for (int i = 0; i < 4; i++) { a.CE[i] = 0; a.CO[i] = 0; }
}
// chain with the carry-out going to a separate counter register
BN_HD void mad_row4_cs(uint32_t* acc, uint32_t& counter, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#if defined(__CUDA_ARCH__)
    BN_MAD_ASM("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(counter)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    uint32_t t[9];
    for (int j = 0; j < 8; j++) t[j] = acc[j];
    t[8] = counter;
    mad_row4(t, x0, x1, x2, x3, y);
    for (int j = 0; j < 8; j++) acc[j] = t[j];
    counter = t[8];
#endif
}
// acc += a * b
BN_HD void acc_mac(AccEO& A, const Fp& a, const Fp& b) {
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        const uint32_t y = b.v[i];
        if ((i & 1) == 0) {
            // even row: E window [i, i+8) -> carry to E position i+8 ; O window [i, i+8) -> carry to O position i+8
            if (i + 8 < 16) {
                mad_row4_cs(&A.E[i], A.CE[i / 2], a.v[0], a.v[2], a.v[4], a.v[6], y);
                mad_row4_cs(&A.O[i], A.CO[i / 2], a.v[1], a.v[3], a.v[5], a.v[7], y);
            }
        } else {
            // odd row: O window [i-1, i+7) -> carry to O position i+7 ; E window [i+1, i+9) -> carry to E position i+9
            mad_row4_cs(&A.O[i - 1], A.CO[(i - 1) / 2], a.v[0], a.v[2], a.v[4], a.v[6], y);
            if (i + 9 < 16)
                mad_row4_cs(&A.E[i + 1], A.CE[(i + 1) / 2], a.v[1], a.v[3], a.v[5], a.v[7], y);
            else
                mad_row4_nc(&A.E[i + 1], a.v[1], a.v[3], a.v[5], a.v[7], y);  // top window: total < 2^512, no carry out
        }
    }
}
// merge to a plain 512-bit integer
BN_HD Wide acc_merge(const AccEO& A) {
    Wide T;
    uint32_t e_hi[8], o_hi[8];
    // fold the counters into the high halves (positions 8,10,12,14)
    const uint32_t ce[8] = {A.CE[0], 0u, A.CE[1], 0u, A.CE[2], 0u, A.CE[3], 0u};
    const uint32_t co[8] = {A.CO[0], 0u, A.CO[1], 0u, A.CO[2], 0u, A.CO[3], 0u};
    (void)add8(e_hi, &A.E[8], ce);
    (void)add8(o_hi, &A.O[8], co);
    BN_UNROLL
    for (int i = 0; i < 8; i++) T.w[i] = A.E[i];
    BN_UNROLL
    for (int i = 0; i < 8; i++) T.w[8 + i] = e_hi[i];
    uint32_t o_all[16];
    BN_UNROLL
    for (int i = 0; i < 8; i++) o_all[i] = A.O[i];
    BN_UNROLL
    for (int i = 0; i < 8; i++) o_all[8 + i] = o_hi[i];
    add16_shift1(T.w, o_all);
    return T;
}

// Montgomery reduction: returns T / 2^256 mod p, in [0, T/2^256 + p).  Caller applies cond_sub.
// Word-serial (HAC 14.32, as reference src/arith.rs:497-500) on the low half only; T's high half is
// added at the end so each chain's carry lands in a fresh limb.  64 IMAD.WIDE + 8 IMAD.
template <class M>
BN_HD void mont_reduce_raw(uint32_t* r, const Wide& T) {
    uint32_t E[18], O[16];
    BN_UNROLL
    for (int i = 0; i < 8; i++) E[i] = T.w[i];
    BN_UNROLL
    for (int i = 8; i < 18; i++) E[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 16; i++) O[i] = 0;
    const uint32_t q0 = M::m(0), q1 = M::m(1), q2 = M::m(2), q3 = M::m(3), q4 = M::m(4), q5 = M::m(5),
                   q6 = M::m(6), q7 = M::m(7);
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        uint32_t low = (i == 0) ? E[0] : (E[i] + O[i - 1]);
        uint32_t mi = low * M::inv();
        if ((i & 1) == 0) {
            mad_row4(&E[i], q0, q2, q4, q6, mi);
            if (i == 0)
                mad_row4(&O[0], q1, q3, q5, q7, mi);
            else
                mad_row4_fold(&O[i], q1, q3, q5, q7, mi, E[i], O[i - 1], true);
        } else {
            mad_row4(&O[i - 1], q0, q2, q4, q6, mi);
            mad_row4_fold(&E[i + 1], q1, q3, q5, q7, mi, E[i], O[i - 1], true);
        }
    }
    // result limbs: positions 8..15 = E[8+j] + O[7+j], then + T_hi
    uint32_t t[8];
    (void)add8(t, &E[8], &O[7]);
    (void)add8(r, t, &T.w[8]);
}

// canonical reduce; MAXK = bound on the raw result in units of p (2 -> one cond-sub, 4 -> two)
template <class M, int MAXK>
BN_HD Fp mont_reduce(const Wide& T) {
    Fp r;
    mont_reduce_raw<M>(r.v, T);
    if (MAXK > 4) {  // < 8p never needed: accumulators are sized so raw < 4p
    }
    if (MAXK > 2) cond_sub_2p<M>(r.v);
    cond_sub_p<M>(r.v);
    return r;
}

// Product and Montgomery reduction interleaved row by row (coarsely integrated operand scanning, the order of reference
// src/arith.rs:257-263 / mul_reduce) on ONE (E, O) pair: row i adds a*b_i (and c*d_i when TWO), then m_i*p with
// m_i = -(limb i)/p mod 2^32, which cancels limb i.  Every chain's carry lands in the window's top limb, which no earlier
// row has multiplied into (as in wide_mac2), so the 512-bit product is never merged: against wide_mac + mont_reduce_raw
// this saves the 16-limb merge of the product and the final addition of its high half (24 IADD3 of ~200 instructions).
// Returns (a*b [+ c*d] + m*p) / 2^256 in [0, (a*b [+ c*d]) / 2^256 + p): the same integer mont_reduce_raw returns.
template <class M, bool TWO>
BN_HD void mul_reduce_rows(uint32_t* r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    uint32_t E[18], O[16];
    BN_UNROLL
    for (int i = 0; i < 18; i++) E[i] = 0;
    BN_UNROLL
    for (int i = 0; i < 16; i++) O[i] = 0;
    const uint32_t q0 = M::m(0), q1 = M::m(1), q2 = M::m(2), q3 = M::m(3), q4 = M::m(4), q5 = M::m(5),
                   q6 = M::m(6), q7 = M::m(7);
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        eo_row(E, O, a.v, b.v[i], i);
        if (TWO) eo_row(E, O, c.v, d.v[i], i);
        uint32_t low = (i == 0) ? E[0] : (E[i] + O[i - 1]);
        uint32_t mi = low * M::inv();
        if ((i & 1) == 0) {
            mad_row4(&E[i], q0, q2, q4, q6, mi);
            if (i == 0)
                mad_row4(&O[0], q1, q3, q5, q7, mi);
            else
                mad_row4_fold(&O[i], q1, q3, q5, q7, mi, E[i], O[i - 1], true);
        } else {
            mad_row4(&O[i - 1], q0, q2, q4, q6, mi);
            mad_row4_fold(&E[i + 1], q1, q3, q5, q7, mi, E[i], O[i - 1], true);
        }
    }
    (void)add8(r, &E[8], &O[7]);
}
// a*b*R^-1 mod p, canonical.   reference src/arith.rs:257-263 (mul_reduce + final correction)
template <class M>
BN_HD Fp fp_mul(const Fp& a, const Fp& b) {
#if BN_MUL_FUSED
    Fp r;
    mul_reduce_rows<M, false>(r.v, a, b, a, b);
    cond_sub_p<M>(r.v);
    return r;
#else
    Wide T = wide_zero();
    wide_mac1(T, a, b);
    return mont_reduce<M, 2>(T);
#endif
}
// (a*b + c*d)*R^-1 mod p, canonical, for a*b + c*d < 2^256 p (e.g. all four < p, or one factor of each product <= p and
// the other < 2p... see the callers)
template <class M>
BN_HD Fp fp_mul2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#if BN_MUL_FUSED
    Fp r;
    mul_reduce_rows<M, true>(r.v, a, b, c, d);
    cond_sub_p<M>(r.v);
    return r;
#else
    Wide T = wide_zero();
    wide_mac2(T, a, b, c, d);
    return mont_reduce<M, 2>(T);
#endif
}
// out-of-line copy for the thread-per-element kernels (keeps their code inside the instruction cache)
template <class M>
BN_HD_NOINLINE Fp fp_mul_ni(Fp a, Fp b) {
    return fp_mul<M>(a, b);
}
// a^2 R^-1 mod p, canonical: dedicated squaring (36 + 72 IMAD against 64 + 72).   reference FieldElement::squared,
// src/fields/mod.rs:31-33 (a * a there)
template <class M>
BN_HD Fp fp_sqr(const Fp& a) {
    Wide T;
    wide_sqr(T, a);
    return mont_reduce<M, 2>(T);
}
template <class M>
BN_HD_NOINLINE Fp fp_sqr_ni(Fp a) {
    return fp_sqr<M>(a);
}
// Montgomery -> plain integer (multiply by 1).   reference src/fields/fp.rs:15-22
template <class M>
BN_HD Fp fp_from_mont(const Fp& a) {
    Wide T = wide_zero();
    BN_UNROLL
    for (int i = 0; i < 8; i++) T.w[i] = a.v[i];
    return mont_reduce<M, 2>(T);
}

// x^(p-2): same canonical value as reference Fq::inverse (src/fields/fp.rs:103-112; binary Euclid there,
// Fermat here -- the inverse is unique, so the Montgomery limbs agree).  x must be non-zero.
template <class M>
BN_HD_NOINLINE Fp fp_inv(Fp x) {
    // exponent p-2, MSB first; p-2 has bit 253 set.
    Fp r = x;
    for (int bit = 252; bit >= 0; bit--) {
        r = fp_sqr<M>(r);
        uint32_t e = (bit < 32) ? (M::m(0) - 2u) : M::m(bit >> 5);  // p-2 only changes limb 0 (p odd, p0 >= 2)
        if ((e >> (bit & 31)) & 1u) r = fp_mul<M>(r, x);
    }
    return r;
}

}  // namespace bn
