// fp64_mac.cuh -- PROTOTYPE (not linked into the product): an exact 256 x 256 -> 512-bit multiply-accumulate on the FP64
// pipe, for DESIGN.md section 9 item 1 (B200: DFMA costs 2.1 issue cycles and overlaps with IMAD.WIDE).
//
// Operands are split into six 44-bit limbs held as doubles (exact integers < 2^44).  For a limb pair p = a*b < 2^88:
//     r  = fma(a, b, C1),  C1 = 1.5 * 2^96      -> r = C1 + p rounded to a multiple of 2^44 (ulp of [2^96, 2^97))
//     hi = r - C1                                  exact, a multiple of 2^44 with |p - hi| <= 2^43
//     lo = fma(a, b, -hi)                          exact (|lo| <= 2^43)
// Column k (weight 2^(44 k)) accumulates lo of the pairs with i + j = k in L[k] and hi in H[k]; with N pairs per column
// the sums need 44 + log2(N) significant bits (H is a multiple of 2^44), exact in a double up to N = 512.  A dense
// Fq12 product accumulates 6 rounds x 6 pairs = 36 pairs per column.
// Cost per 6 x 6 product: 36 x (2 DFMA + 3 DADD) = 180 FP64 instructions, no integer instruction.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FP64_HD __host__ __device__ __forceinline__
#else
#define FP64_HD inline
#endif

namespace fp64proto {

struct D6 {
    double l[6];
};
struct Cols {
    double H[11], L[11];
};

FP64_HD void cols_zero(Cols& c) {
    for (int k = 0; k < 11; k++) c.H[k] = c.L[k] = 0.0;
}

// 8 x 32-bit limbs (value < 2^256) -> 6 x 44-bit limbs as doubles.  Integer side: per limb two funnel shifts and a mask,
// then the 2^52 bit-pattern trick (no conversion instruction) and one DADD.
FP64_HD D6 to_d6(const uint32_t* v) {
    D6 r;
    for (int k = 0; k < 6; k++) {
        const int bit = 44 * k, w = bit >> 5, s = bit & 31;
        uint64_t lo = v[w];
        uint64_t mid = (w + 1 < 8) ? v[w + 1] : 0u;
        uint64_t hi = (w + 2 < 8) ? v[w + 2] : 0u;
        uint64_t x = (lo >> s) | (mid << (32 - s)) | (s ? (hi << (64 - s)) : 0u);
        x &= ((uint64_t)1 << 44) - 1;
        uint64_t pat = 0x4330000000000000ULL | x;  // 2^52 + x
        double d;
#if defined(__CUDA_ARCH__)
        d = __longlong_as_double((long long)pat);
#else
        __builtin_memcpy(&d, &pat, 8);
#endif
        r.l[k] = d - 4503599627370496.0;  // - 2^52
    }
    return r;
}

// C += a * b  (exact)
FP64_HD void mac(Cols& c, const D6& a, const D6& b) {
    const double C1 = 1.5 * 79228162514264337593543950336.0;  // 1.5 * 2^96
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            const double r = fma(a.l[i], b.l[j], C1);
            const double hi = r - C1;
            const double lo = fma(a.l[i], b.l[j], -hi);
            c.H[i + j] += hi;
            c.L[i + j] += lo;
        }
}

// sum_k (L[k] + H[k]) 2^(44 k)  mod 2^544, as 17 x 32-bit limbs (two's complement; the caller takes the low 512 bits).
// Integer side, once per operation: column value c_k = L[k] + H[k-1] / 2^44 is an exact integer with |c_k| < 2^53.
FP64_HD void cols_to_limbs(const Cols& c, uint32_t* out17) {
    for (int i = 0; i < 17; i++) out17[i] = 0;
    long long carry = 0;  // running signed carry in units of the current column
    // walk bit positions: column k starts at bit 44 k
    unsigned __int128 acc = 0;  // host/device-neutral enough for the prototype (device: two 64-bit words)
    int accbits = 0;            // bits of acc already emitted
    (void)acc; (void)accbits;
    // simple exact evaluation with signed 128-bit arithmetic per column, emitted limb by limb
    __int128 run = 0;  // value of everything below the current emission point, shifted down
    int emitted = 0;   // number of bits already written to out17
    for (int k = 0; k <= 11; k++) {
        __int128 ck = 0;
        if (k < 11) ck += (__int128)(long long)c.L[k];
        if (k >= 1) ck += (__int128)(long long)(c.H[k - 1] * (1.0 / 17592186044416.0));  // / 2^44, exact
        // run holds bits from position `emitted`; column k sits at bit 44 k
        run += ck << (44 * k - emitted);
        // emit all complete 32-bit limbs below the next column's start
        const int next = 44 * (k + 1);
        while (emitted + 32 <= next && emitted < 544) {
            out17[emitted >> 5] = (uint32_t)(run & 0xffffffffu);
            run >>= 32;  // arithmetic shift keeps the sign
            emitted += 32;
        }
    }
    while (emitted < 544) {
        out17[emitted >> 5] = (uint32_t)(run & 0xffffffffu);
        run >>= 32;
        emitted += 32;
    }
    (void)carry;
}

}  // namespace fp64proto
