// f52.cuh -- 256-bit modular products on the FP64 pipe: 5 x 52-bit limbs held as doubles.
//
// Why: on B200 a 32x32+64 IMAD.WIDE costs the scheduler 4 cycles per warp instruction (8 lanes/clk on fmaheavy), and a
// 256x256 product needs 64 of them (256 cycles) plus ~32 carry/merge instructions.  The FP64 pipe issues a DFMA every
// ~2 cycles (16 lanes/clk, B200 keeps full-rate FP64) and one DFMA pair yields an EXACT 52x52 -> 104-bit product:
//     hi = fma_rz(a, b, 2^104)                  -> bit pattern 0x467 | floor(a b / 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)    -> bit pattern 0x433 | (a b mod 2^52)
// (round toward zero; a, b integers in [0, 2^52)).  25 limb products = 50 DFMA + 25 DADD = ~150 FP64-pipe cycles, and
// the 50 partial words are summed as 64-bit INTEGER bit patterns by 25 three-input adds (IADD3 + IADD3.X) on the ALU
// pipe, which runs concurrently.  The exponent patterns are compile-time constants removed once per accumulator.
// Measured (tools/ubench/f52.cu, profiles/r02_run2_ubench_f52.txt): products 1.5x cheaper than fp.cuh's IMAD.WIDE rows even
// in an untuned chain kernel; results exact against host big integers.
//
// Scope: only the hot inner operation of the hexad kernels -- the lazy Fq2 multiply-accumulate and its Montgomery
// reduction.  Everything between those (additions, xi-multiplication, conjugation, parking, I/O) stays on canonical
// 8 x u32 Montgomery values (fp.cuh / fp2.cuh): a lane converts an operand to doubles once, when it publishes it to the
// hexad's shared-memory slots, and the reduction repacks its result into 8 x u32.  The Montgomery radix stays 2^256
// (four 52-bit rounds and one 48-bit round), so every constant, table and byte layout of the integer code is unchanged
// and the results are the same canonical values, bit for bit.
#pragma once
#include "fp2.cuh"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

namespace bn {

typedef unsigned long long u64;
typedef long long i64;

struct D5 {  // value = sum l[i] 2^(52 i), every l[i] an integer in [0, 2^52)
    double l[5];
};
// an Fq2 operand as the multiply-accumulate consumes it: c0, c1 and c0 + c1 (Karatsuba's third factor, summed by the
// publisher in the integer domain so the limbs stay normalised)
struct D5x3 {
    D5 c0, c1, cs;
};

#define F52_MASK 0x000FFFFFFFFFFFFFULL
#define F52_MASK48 0x0000FFFFFFFFFFFFULL
#define F52_LO_OFF 0x4330000000000000ULL  // bit pattern of 2^52
#define F52_HI_OFF 0x4670000000000000ULL  // bit pattern of 2^104
#define F52_TWO52 4503599627370496.0

BN_HD double f52_from_bits(u64 b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
BN_HD u64 f52_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (u64)__double_as_longlong(d);
#else
    u64 b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
// fma with round-toward-zero.  Host builds (tests/host_emu) run with the FPU in FE_TOWARDZERO mode (set per thread by
// the emulator), so a plain fma is the same operation there.
BN_HD double f52_fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

// exact 52 x 52 -> 104-bit product as two bit patterns: hi = F52_HI_OFF | floor(ab / 2^52), lo = F52_LO_OFF | (ab mod 2^52)
BN_HD void f52_limb_mul(double a, double b, u64& hi, u64& lo) {
    const double c1 = f52_from_bits(F52_HI_OFF), c2 = f52_from_bits(F52_HI_OFF + 1);
    const double ph = f52_fma_rz(a, b, c1);
    const double sb = c2 - ph;  // exact
    const double pl = f52_fma_rz(a, b, sb);
    hi = f52_bits(ph);
    lo = f52_bits(pl);
}

// integer < 2^52 given as (lo32, hi20) -> double, without a conversion instruction
BN_HD double f52_make(uint32_t lo, uint32_t hi20) { return f52_from_bits(((u64)(hi20 | 0x43300000u) << 32) | lo) - F52_TWO52; }

// 8 x u32 (any integer < 2^256) -> 5 normalised 52-bit limbs.  ~13 integer instructions + 5 DADD.
BN_HD D5 f52_from_fp(const Fp& a) {
    D5 r;
    const uint32_t* v = a.v;
    r.l[0] = f52_make(v[0], v[1] & 0xFFFFFu);
    r.l[1] = f52_make((v[1] >> 20) | (v[2] << 12), ((v[2] >> 20) | (v[3] << 12)) & 0xFFFFFu);
    r.l[2] = f52_make((v[3] >> 8) | (v[4] << 24), (v[4] >> 8) & 0xFFFFFu);
    r.l[3] = f52_make((v[4] >> 28) | (v[5] << 4), ((v[5] >> 28) | (v[6] << 4)) & 0xFFFFFu);
    r.l[4] = f52_make((v[6] >> 16) | (v[7] << 16), v[7] >> 16);
    return r;
}
// an Fq2 value with components < 2q (canonical or a raw sum of two canonical values) -> the three factors
BN_HD D5x3 f52_from_fp2(const Fp2& a) {
    D5x3 r;
    r.c0 = f52_from_fp(a.c0);
    r.c1 = f52_from_fp(a.c1);
    r.cs = f52_from_fp(fp_add_raw(a.c0, a.c1));  // < 4q < 2^256
    return r;
}

// ------------------------------------------------------------------------------------------------
// Column accumulators.  Column k has weight 2^(52 k); it collects the lo patterns of the limb pairs with i + j = k and
// the hi patterns of the pairs with i + j = k - 1, as plain 64-bit integer sums (mod 2^64; the true column values are
// far below 2^63: at most 10 terms of 52 bits per accumulated product).
// ------------------------------------------------------------------------------------------------
struct Cols {
    i64 c[11];
};
BN_HD void cols_zero(Cols& C) {
    BN_UNROLL
    for (int k = 0; k < 11; k++) C.c[k] = 0;
}
// number of lo / hi patterns a single 5 x 5 product puts into column k
BN_HD int f52_nlo(int k) { return k <= 4 ? k + 1 : (k <= 8 ? 9 - k : 0); }
BN_HD int f52_nhi(int k) { return k == 0 ? 0 : f52_nlo(k - 1); }
// the exponent-pattern total of `n` accumulated products in column k (mod 2^64)
BN_HD u64 f52_offset(int k, int n) { return (u64)n * ((u64)f52_nlo(k) * F52_LO_OFF + (u64)f52_nhi(k) * F52_HI_OFF); }

// C += a * b   (75 FP64 instructions, 25 three-input 64-bit adds)
BN_HD void f52_mac(Cols& C, const D5& a, const D5& b) {
    BN_UNROLL
    for (int i = 0; i < 5; i++) {
        BN_UNROLL
        for (int j = 0; j < 5; j++) {
            u64 h, l;
            f52_limb_mul(a.l[i], b.l[j], h, l);
            C.c[i + j] += (i64)l;
            C.c[i + j + 1] += (i64)h;
        }
    }
}

#include "constants_f52.inc"

// Montgomery reduction.  On entry the columns hold (true value T) + (the pattern offsets of `nprod` products, already
// compensated by the caller except for ONE product's worth, which this function's own q-multiples re-create) -- see
// f52_reduce's callers: they pass columns whose offsets sum to exactly -1 product.  0 <= T < 2^512 - q 2^256.
// Returns T / 2^256 mod q as 8 x u32, in [0, T / 2^256 + q): the caller applies the conditional subtractions.
BN_HD void f52_reduce_raw(uint32_t* out, Cols& C) {
    BN_UNROLL
    for (int i = 0; i < 5; i++) {
        const u64 mask = (i < 4) ? F52_MASK : F52_MASK48;  // 4 x 52 + 48 = 256
        // m = (column i mod 2^w) * (-q^-1) mod 2^w
        const double d = f52_from_bits(((u64)C.c[i] & mask) | F52_LO_OFF) - F52_TWO52;
        u64 h, l;
        f52_limb_mul(d, F52_QINV, h, l);
        const double m = f52_from_bits((l & mask) | F52_LO_OFF) - F52_TWO52;
        BN_UNROLL
        for (int j = 0; j < 5; j++) {
            f52_limb_mul(m, F52_Q_f(j), h, l);
            C.c[i + j] += (i64)l;
            C.c[i + j + 1] += (i64)h;
        }
        if (i < 4) C.c[i + 1] += C.c[i] >> 52;  // column i is now a multiple of 2^52: its carry moves up
    }
    // columns 4..9 (+ carries) hold T' = T + M q, a multiple of 2^256; result = T' >> 256 = stream >> 48 where
    // stream = sum_{k>=4} n_k 2^(52 (k-4)) with normalised n_k
    u64 n[7];
    BN_UNROLL
    for (int k = 4; k < 10; k++) {
        n[k - 4] = (u64)C.c[k] & F52_MASK;
        C.c[k + 1] += C.c[k] >> 52;
    }
    n[6] = (u64)C.c[10];  // zero for in-range inputs
    BN_UNROLL
    for (int t = 0; t < 8; t++) {
        const int pos = 48 + 32 * t, k = pos / 52, off = pos % 52;
        u64 v = n[k] >> off;
        if (off > 20) v |= n[k + 1] << (52 - off);
        out[t] = (uint32_t)v;
    }
}

// ------------------------------------------------------------------------------------------------
// Lazy Fq2 multiply-accumulate (the FP64 counterpart of hexad.cuh's AccK2S): three column sets
//   S0 = sum x0 y0,  S1 = sum x1 y1,  S2 = sum (x0+x1)(y0+y1)
// recombined once per operation into  re = 6 q^2 + S0 - S1,  im = S2 - S0 - S1  and reduced.  Same bounds as the integer
// accumulators: operands < 2q per component, at most six accumulated Fq2 products (re, im < 12 q^2).
// ------------------------------------------------------------------------------------------------
struct Acc52 {
    Cols s0, s1, s2;
};
BN_HD void acc52_init(Acc52& A) {
    cols_zero(A.s0);
    cols_zero(A.s1);
    cols_zero(A.s2);
}
BN_HD void acc52_mac(Acc52& A, const D5x3& x, const D5x3& y) {
    f52_mac(A.s0, x.c0, y.c0);
    f52_mac(A.s1, x.c1, y.c1);
    f52_mac(A.s2, x.cs, y.cs);
}
// one shared copy of the two reductions (code footprint); the 2 x 11 columns travel in registers
BN_HD_NOINLINE Fp2 f52_reduce2(Cols re, Cols im) {
    Fp2 r;
    f52_reduce_raw(r.c0.v, re);
    f52_reduce_raw(r.c1.v, im);
    cond_sub_2p<MQ>(r.c0.v);
    cond_sub_p<MQ>(r.c0.v);
    cond_sub_2p<MQ>(r.c1.v);
    cond_sub_p<MQ>(r.c1.v);
    return r;
}
// NPROD = number of Fq2 products accumulated (compile-time: it fixes the pattern offsets)
template <int NPROD>
BN_HD Fp2 acc52_reduce(const Acc52& A) {
    Cols re, im;
    BN_UNROLL
    for (int k = 0; k < 11; k++) {
        // re: the offsets of S0 and S1 cancel; add 6 q^2 and pre-compensate the reduction's own product (-1)
        re.c[k] = A.s0.c[k] - A.s1.c[k] + (i64)(F52_6Q2_f(k) - f52_offset(k, 1));
        // im: S2 - S0 - S1 leaves -NPROD offsets; bring it to -1 (the reduction's own)
        im.c[k] = A.s2.c[k] - A.s0.c[k] - A.s1.c[k] - (i64)f52_offset(k, 1 - NPROD);
    }
    return f52_reduce2(re, im);
}

// single Fq product through the FP64 path (tests and the thread-level helpers): a b 2^-256 mod q, canonical
BN_HD Fp f52_fp_mul(const Fp& a, const Fp& b) {
    Cols C;
    BN_UNROLL
    for (int k = 0; k < 11; k++) C.c[k] = -(i64)f52_offset(k, 2);  // this product + the reduction's
    f52_mac(C, f52_from_fp(a), f52_from_fp(b));
    Fp r;
    f52_reduce_raw(r.v, C);
    cond_sub_p<MQ>(r.v);
    return r;
}

}  // namespace bn
