// f52.cu -- feasibility micro-benchmark: 256-bit Montgomery multiplication on the FP64 pipe (5 x 52-bit limbs held as
// doubles, products split exactly into high / low halves by two round-toward-zero DFMAs, partial products summed as 64-bit
// integer bit patterns) against the IMAD.WIDE implementation of bn_b200/csrc/fp.cuh, on B200 (sm_100a).
//
//   per 52 x 52 limb product:  hi = fma_rz(a, b, 2^104)            mantissa = floor(a b / 2^52)
//                              lo = fma_rz(a, b, (2^104 + 2^52) - hi)   mantissa = a b mod 2^52
//   column k (weight 2^(52 k)) += pattern(lo_ij), i + j = k;  column k + 1 += pattern(hi_ij); the exponent patterns are
//   compile-time constants folded into the columns' initial values.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o f52 f52.cu
// Every kernel's result is checked against unsigned __int128 host arithmetic before its time is printed.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../bn_b200/csrc/fp.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

typedef unsigned long long u64;
typedef unsigned __int128 u128;

// ---- host big-int helpers (plain schoolbook, 64-bit words) -----------------------------------------------------------
static const u64 Q64[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
struct Big {  // up to 640 bits
    u64 w[10];
};
static Big big_zero() { Big r; memset(&r, 0, sizeof r); return r; }
static Big big_mul(const Big& a, const Big& b, int na, int nb) {
    Big r = big_zero();
    for (int i = 0; i < na; i++) {
        u128 c = 0;
        for (int j = 0; j < nb && i + j < 10; j++) {
            c += (u128)a.w[i] * b.w[j] + r.w[i + j];
            r.w[i + j] = (u64)c;
            c >>= 64;
        }
        if (i + nb < 10) r.w[i + nb] += (u64)c;
    }
    return r;
}
static int big_cmp(const Big& a, const Big& b) {
    for (int i = 9; i >= 0; i--)
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}
static Big big_sub(const Big& a, const Big& b) {
    Big r;
    u128 br = 0;
    for (int i = 0; i < 10; i++) {
        u128 t = (u128)a.w[i] - b.w[i] - br;
        r.w[i] = (u64)t;
        br = (t >> 64) & 1;
    }
    return r;
}
static Big big_add(const Big& a, const Big& b) {
    Big r;
    u128 c = 0;
    for (int i = 0; i < 10; i++) {
        c += (u128)a.w[i] + b.w[i];
        r.w[i] = (u64)c;
        c >>= 64;
    }
    return r;
}
static Big big_shr(const Big& a, int bits) {
    Big r = big_zero();
    int ws = bits / 64, bs = bits % 64;
    for (int i = 0; i + ws < 10; i++) {
        r.w[i] = a.w[i + ws] >> bs;
        if (bs && i + ws + 1 < 10) r.w[i] |= a.w[i + ws + 1] << (64 - bs);
    }
    return r;
}
static Big big_q() { Big r = big_zero(); memcpy(r.w, Q64, 32); return r; }
// x mod q by shift-subtract (x < 2^600)
static Big big_mod_q(Big x) {
    Big q = big_q();
    for (int sh = 340; sh >= 0; sh--) {
        Big t = big_zero();  // q << sh
        int ws = sh / 64, bs = sh % 64;
        for (int i = 0; i < 4; i++) {
            if (i + ws < 10) t.w[i + ws] |= q.w[i] << bs;
            if (bs && i + ws + 1 < 10) t.w[i + ws + 1] |= q.w[i] >> (64 - bs);
        }
        if (big_cmp(x, t) >= 0) x = big_sub(x, t);
    }
    return x;
}
// Montgomery product with R = 2^rbits: a b R^-1 mod q, canonical.  Uses q' = -q^-1 mod 2^64 word-serially after aligning.
static u64 inv64() {
    u64 x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - Q64[0] * x;
    return (u64)0 - x;
}
static Big mont_host(const Big& a, const Big& b, int rbits) {
    // T = a b; add multiples of q to clear rbits low bits, bit by bit in 52/64-bit independent way: do it 4 bits at a time
    Big T = big_mul(a, b, 5, 5);
    Big q = big_q();
    const u64 qi = inv64();
    int done = 0;
    while (done < rbits) {
        int step = std::min(32, rbits - done);
        // low `step` bits of (T >> done)
        Big s = big_shr(T, done);
        u64 low = s.w[0] & ((step == 64) ? ~0ULL : ((1ULL << step) - 1));
        u64 m = (low * qi) & ((1ULL << step) - 1);
        Big mq = big_zero();
        mq.w[0] = m;
        mq = big_mul(mq, q, 1, 4);
        // T += mq << done
        Big t = big_zero();
        int ws = done / 64, bs = done % 64;
        for (int i = 0; i < 6; i++) {
            if (i + ws < 10) t.w[i + ws] |= mq.w[i] << bs;
            if (bs && i + ws + 1 < 10) t.w[i + ws + 1] |= mq.w[i] >> (64 - bs);
        }
        T = big_add(T, t);
        done += step;
    }
    return big_mod_q(big_shr(T, rbits));
}

// ---- F52: device side ------------------------------------------------------------------------------------------------
#define F52_MASK ((1ULL << 52) - 1)
struct F52 {
    double l[5];
};
__host__ __device__ inline double bits_to_double(u64 b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
// modulus limbs (52-bit) and -q^-1 mod 2^52, as doubles
__constant__ double c_q52[5];
__constant__ double c_qinv52;

#define C1_BITS 0x4670000000000000ULL  // 2^104
#define C2_BITS 0x4670000000000001ULL  // 2^104 + 2^52
#define LO_OFF 0x4330000000000000ULL   // pattern of 2^52
#define HI_OFF 0x4670000000000000ULL

__device__ __forceinline__ void limb_mul(double a, double b, u64& hi, u64& lo) {
    const double c1 = __longlong_as_double((long long)C1_BITS), c2 = __longlong_as_double((long long)C2_BITS);
    double ph = __fma_rz(a, b, c1);
    double sb = c2 - ph;  // exact
    double pl = __fma_rz(a, b, sb);
    hi = (u64)__double_as_longlong(ph);
    lo = (u64)__double_as_longlong(pl);
}

// C[0..10] += a * b (columns of bit patterns; the caller pre-loaded the offset corrections)
__device__ __forceinline__ void f52_mac(u64* C, const F52& a, const F52& b) {
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            u64 h, l;
            limb_mul(a.l[i], b.l[j], h, l);
            C[i + j] += l;
            C[i + j + 1] += h;
        }
}
// offset correction for NP accumulated products: column k holds n_lo(k) lo patterns and n_hi(k) hi patterns
template <int NP>
__device__ __forceinline__ void f52_cols_init(u64* C) {
#pragma unroll
    for (int k = 0; k < 11; k++) {
        int nlo = 0, nhi = 0;
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 5; j++) {
                if (i + j == k) nlo++;
                if (i + j + 1 == k) nhi++;
            }
        C[k] = (u64)0 - (u64)NP * ((u64)nlo * LO_OFF + (u64)nhi * HI_OFF);
    }
}
// Montgomery reduction of the columns (true value T < q 2^260): returns T / 2^260 mod q in [0, 2q), normalised limbs
__device__ __forceinline__ F52 f52_reduce(u64* C) {
    const double two52 = 4503599627370496.0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        // low 52 bits of column i as a double, m = low * qinv mod 2^52
        double d = __longlong_as_double((long long)((C[i] & F52_MASK) | LO_OFF)) - two52;
        u64 h, l;
        limb_mul(d, c_qinv52, h, l);
        double m = __longlong_as_double((long long)l) - two52;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            limb_mul(m, c_q52[j], h, l);
            C[i + j] += l - LO_OFF;
            C[i + j + 1] += h - HI_OFF;
        }
        C[i + 1] += C[i] >> 52;  // column i is now a multiple of 2^52
    }
    F52 r;
#pragma unroll
    for (int k = 5; k < 10; k++) {
        r.l[k - 5] = __longlong_as_double((long long)((C[k] & F52_MASK) | LO_OFF)) - two52;
        C[k + 1] += C[k] >> 52;
    }
    return r;
}
__device__ __forceinline__ F52 f52_mul(const F52& a, const F52& b) {
    u64 C[11];
    f52_cols_init<1>(C);
    f52_mac(C, a, b);
    return f52_reduce(C);
}

// chain x <- x * y (full Montgomery multiplications)
__global__ void __launch_bounds__(256) k_f52_chain(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F52 x, y;
#pragma unroll
    for (int k = 0; k < 5; k++) { x.l[k] = a[i * 5 + k]; y.l[k] = b[i * 5 + k]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) x = f52_mul(x, y);
#pragma unroll
    for (int k = 0; k < 5; k++) out[i * 5 + k] = x.l[k];
}
// lazy form: x <- (x*y + x*z + y*z + x*x) / R : four products per reduction (what the tower code does)
__global__ void __launch_bounds__(256) k_f52_lazy4(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F52 x, y, z;
#pragma unroll
    for (int k = 0; k < 5; k++) { x.l[k] = a[i * 5 + k]; y.l[k] = b[i * 5 + k]; }
    z = f52_mul(y, y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        u64 C[11];
        f52_cols_init<4>(C);
        f52_mac(C, x, y);
        f52_mac(C, x, z);
        f52_mac(C, y, z);
        f52_mac(C, x, x);
        x = f52_reduce(C);  // T < 4 (2q)^2 = 16 q^2 < q 2^260: result < 2q
    }
#pragma unroll
    for (int k = 0; k < 5; k++) out[i * 5 + k] = x.l[k];
}

// ---- the IMAD.WIDE implementation of the product (bn_b200/csrc/fp.cuh), same two shapes --------------------------------
using namespace bn;
__global__ void __launch_bounds__(256) k_imad_chain(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x, y;
#pragma unroll
    for (int k = 0; k < 8; k++) { x.v[k] = a[i * 8 + k]; y.v[k] = b[i * 8 + k]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) x = fp_mul<ModQ>(x, y);
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 8 + k] = x.v[k];
}
__global__ void __launch_bounds__(256) k_imad_lazy4(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x, y, z;
#pragma unroll
    for (int k = 0; k < 8; k++) { x.v[k] = a[i * 8 + k]; y.v[k] = b[i * 8 + k]; }
    z = fp_mul<ModQ>(y, y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        Wide T = wide_zero();
        wide_mac2(T, x, y, x, z);
        wide_mac2(T, y, z, x, x);
        x = mont_reduce<ModQ, 4>(T);  // canonical
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 8 + k] = x.v[k];
}

// ---- pipe-rate probes: cycles per warp instruction at W warps per scheduler ---------------------------------------------
// kind 0: DFMA.RZ on register operands (8 independent chains)      kind 1: DADD
// kind 2: plain IMAD.WIDE.U32 (no carry), 8 independent accumulators   kind 3: IMAD.LO (32-bit)   kind 4: IMAD.HI
// kind 5: IMAD.WIDE.U32.X chains of 4 (as in fp.cuh)                kind 6: 64-bit integer add of two values (3-input)
// kind 7: DFMA + 64-bit 3-input adds in the limb_mul shape (2 DFMA + 1 DADD + 2 adds per product)
template <int KIND>
__global__ void k_pipe(u64* out, long long* cyc, int iters, double seed, unsigned useed) {
    double d[8], x = seed * 1.0000001 + threadIdx.x, y = seed * 0.9999999;
    u64 m[8];
    unsigned a[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = seed + i + threadIdx.x; m[i] = (u64)i * 0x9E3779B97F4A7C15ULL + threadIdx.x + useed; a[i] = useed * (2 * i + 3) + threadIdx.x; lo[i] = a[i] ^ 0x5555u; }
    unsigned b = threadIdx.x * 40503u + 7u + useed;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (KIND == 0) d[i] = __fma_rz(d[i], x, y);
                if (KIND == 1) d[i] = __dadd_rz(d[i], x);
                if (KIND == 2) m[i] = (u64)a[i] * (u64)(b + r) + m[i];
                if (KIND == 3) lo[i] = a[i] * (b + r) + lo[i];
                if (KIND == 4) lo[i] = __umulhi(a[i], lo[i] | 1u) + b;
                if (KIND == 6) m[i] = m[i] + m[(i + 1) & 7] + m[(i + 3) & 7];
                if (KIND == 7) {
                    u64 h, l;
                    limb_mul(d[i], x, h, l);
                    m[i] += h + l;
                    d[i] = __longlong_as_double((long long)((l & F52_MASK) | LO_OFF)) - 4503599627370496.0;
                }
            }
            if (KIND == 5) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    uint32_t* acc = reinterpret_cast<uint32_t*>(&m[4 * j]);
                    mad_row4_nc(acc, a[0], a[1], a[2], a[3], b + r + j);
                    mad_row4_nc(acc, a[4], a[5], a[6], a[7], b + r + j);
                    mad_row4_nc(acc, a[0], a[1], a[2], a[3], lo[j] + r);
                    mad_row4_nc(acc, a[4], a[5], a[6], a[7], lo[j + 2] + r);
                }
            }
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= m[i] ^ (u64)__double_as_longlong(d[i]) ^ lo[i];
    if (s == 0x123456789ULL) out[0] = s;
}

static int n_sm;
static u64* d_out;
static long long* d_cyc;
template <int KIND>
static double pipe_cycles(int warps_per_smsp, int iters) {
    int threads = 128 * warps_per_smsp, blocks = 1;
    if (threads > 1024) { blocks = threads / 512; threads = 512; }
    int grid = n_sm * blocks;
    for (int rep = 0; rep < 2; rep++) {
        k_pipe<KIND><<<grid, threads>>>(d_out, d_cyc, iters, 1.0 + 1e-9, 12345u);
        cudaDeviceSynchronize();
    }
    int nw = grid * threads / 32;
    std::vector<long long> h(nw);
    cudaMemcpy(h.data(), d_cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    return (double)h[nw / 2] / iters / 32.0;  // per unrolled instruction group member (32 per trip)
}

static void to52(const u64* w4, double* l5) {
    // 256-bit value (4 x u64) -> 5 x 52-bit limbs
    Big b = big_zero();
    memcpy(b.w, w4, 32);
    for (int k = 0; k < 5; k++) {
        Big s = big_shr(b, 52 * k);
        l5[k] = (double)(s.w[0] & F52_MASK);
    }
}
static Big from52(const double* l5) {
    Big r = big_zero();
    for (int k = 4; k >= 0; k--) {
        // r = (r << 52) + limb
        Big t = big_zero();
        for (int i = 0; i < 9; i++) {
            t.w[i] |= r.w[i] << 52;
            t.w[i + 1] |= r.w[i] >> 12;
        }
        Big l = big_zero();
        l.w[0] = (u64)l5[k];
        r = big_add(t, l);
    }
    return r;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    n_sm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, n_sm, p.clockRate);
    CK(cudaMalloc(&d_out, 4096));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * n_sm * 64 * 8));

    // constants
    double q52[5];
    to52(Q64, q52);
    u64 qi = inv64() & F52_MASK;
    double qinv = (double)qi;
    CK(cudaMemcpyToSymbol(c_q52, q52, sizeof q52));
    CK(cudaMemcpyToSymbol(c_qinv52, &qinv, sizeof qinv));

    // ---- pipe probes
    printf("\n[P] cycles per warp instruction (median warp), W warps per scheduler, one or more blocks per SM\n");
    printf("%-44s %7s %7s %7s %7s %7s %7s\n", "instruction", "W=1", "W=2", "W=3", "W=4", "W=8", "W=16");
    const int it = 2000;
#define PROW(KIND, NAME, DIV)                                                                                       \
    {                                                                                                               \
        printf("%-44s", NAME);                                                                                      \
        int ws[6] = {1, 2, 3, 4, 8, 16};                                                                            \
        for (int w = 0; w < 6; w++) printf(" %7.2f", pipe_cycles<KIND>(ws[w], it) / (DIV) * ws[w] / ws[w]);         \
        printf("   (per warp; divide by W for the scheduler's rate)\n");                                            \
    }
    PROW(0, "DFMA.RZ (register operands)", 1.0)
    PROW(1, "DADD.RZ", 1.0)
    PROW(2, "IMAD.WIDE.U32 (no carry, 8 independent)", 1.0)
    PROW(3, "IMAD (32-bit lo)", 1.0)
    PROW(4, "IMAD.HI.U32", 1.0)
    PROW(5, "IMAD.WIDE.U32.X (4-chains, fp.cuh rows) /IMAD", 1.0)
    PROW(6, "64-bit add, three inputs (IADD3 + IADD3.X)", 1.0)
    PROW(7, "limb_mul + accumulate (2 DFMA, 2 DADD, ints)", 1.0)

    // ---- multiplication chains: correctness, then throughput
    const size_t n = (size_t)n_sm * 2048 * 4;
    const int iters = 512;
    std::vector<u64> ha(n * 4), hb(n * 4);
    u64 st = 0xB2000002ULL;
    auto next = [&]() { st += 0x9E3779B97F4A7C15ULL; u64 z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); };
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 4; k++) { ha[i * 4 + k] = next(); hb[i * 4 + k] = next(); }
        ha[i * 4 + 3] &= 0x0FFFFFFFFFFFFFFFULL;  // < q
        hb[i * 4 + 3] &= 0x0FFFFFFFFFFFFFFFULL;
    }
    std::vector<double> fa(n * 5), fb(n * 5), fo(n * 5);
    for (size_t i = 0; i < n; i++) { to52(&ha[i * 4], &fa[i * 5]); to52(&hb[i * 4], &fb[i * 5]); }
    double *dfa, *dfb, *dfo;
    uint32_t *dia, *dib, *dio;
    CK(cudaMalloc(&dfa, n * 40)); CK(cudaMalloc(&dfb, n * 40)); CK(cudaMalloc(&dfo, n * 40));
    CK(cudaMalloc(&dia, n * 32)); CK(cudaMalloc(&dib, n * 32)); CK(cudaMalloc(&dio, n * 32));
    CK(cudaMemcpy(dfa, fa.data(), n * 40, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dfb, fb.data(), n * 40, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dia, ha.data(), n * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dib, hb.data(), n * 32, cudaMemcpyHostToDevice));

    // correctness of f52_mul: 3 chained multiplications on a few elements against the host (R = 2^260)
    {
        k_f52_chain<<<(unsigned)((n + 255) / 256), 256>>>(dfa, dfb, dfo, n, 3);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(fo.data(), dfo, n * 40, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (size_t i = 0; i < 64; i++) {
            Big x = big_zero(), y = big_zero();
            memcpy(x.w, &ha[i * 4], 32);
            memcpy(y.w, &hb[i * 4], 32);
            for (int t = 0; t < 3; t++) x = mont_host(x, y, 260);
            Big got = big_mod_q(from52(&fo[i * 5]));
            bool limbs_ok = true;
            for (int k = 0; k < 5; k++) limbs_ok = limbs_ok && fo[i * 5 + k] >= 0 && fo[i * 5 + k] < 4503599627370496.0;
            if (big_cmp(got, x) != 0 || !limbs_ok) bad++;
        }
        printf("\n[C] f52_mul (5 x 52-bit limbs, R = 2^260): %s on 64 elements x 3 chained multiplications vs host big integers\n", bad ? "MISMATCH" : "exact");
        k_f52_lazy4<<<(unsigned)((n + 255) / 256), 256>>>(dfa, dfb, dfo, n, 2);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(fo.data(), dfo, n * 40, cudaMemcpyDeviceToHost));
        bad = 0;
        for (size_t i = 0; i < 64; i++) {
            Big x = big_zero(), y = big_zero();
            memcpy(x.w, &ha[i * 4], 32);
            memcpy(y.w, &hb[i * 4], 32);
            Big z = mont_host(y, y, 260);
            for (int t = 0; t < 2; t++) {
                // (xy + xz + yz + xx) / R mod q : linear, so sum of the four Montgomery products mod q
                Big s = big_add(big_add(mont_host(x, y, 260), mont_host(x, z, 260)), big_add(mont_host(y, z, 260), mont_host(x, x, 260)));
                x = big_mod_q(s);
            }
            if (big_cmp(big_mod_q(from52(&fo[i * 5])), x) != 0) bad++;
        }
        printf("[C] f52 lazy accumulation (4 products per reduction): %s\n", bad ? "MISMATCH" : "exact");
    }

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time_ms = [&](auto launch) {
        launch();
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        return (double)best;
    };
    const unsigned grid = (unsigned)((n + 255) / 256);
    double t_f = time_ms([&] { k_f52_chain<<<grid, 256>>>(dfa, dfb, dfo, n, iters); });
    double t_i = time_ms([&] { k_imad_chain<<<grid, 256>>>(dia, dib, dio, n, iters); });
    double t_f4 = time_ms([&] { k_f52_lazy4<<<grid, 256>>>(dfa, dfb, dfo, n, iters); });
    double t_i4 = time_ms([&] { k_imad_lazy4<<<grid, 256>>>(dia, dib, dio, n, iters); });
    printf("\n[T] %zu lanes x %d iterations\n", n, iters);
    printf("  full Montgomery multiplication   FP64 (f52): %8.3f ms  %.3e mul/s    IMAD.WIDE (fp.cuh): %8.3f ms  %.3e mul/s    ratio %.2f\n",
           t_f, n * (double)iters / (t_f * 1e-3), t_i, n * (double)iters / (t_i * 1e-3), t_i / t_f);
    printf("  4 products + 1 reduction         FP64 (f52): %8.3f ms  %.3e prod/s   IMAD.WIDE (fp.cuh): %8.3f ms  %.3e prod/s   ratio %.2f\n",
           t_f4, 4.0 * n * iters / (t_f4 * 1e-3), t_i4, 4.0 * n * iters / (t_i4 * 1e-3), t_i4 / t_f4);
    // derived: cost of one product and of one reduction in each implementation (ms per n*iters)
    double pf = (t_f4 - t_f) / 3, rf = t_f - pf, pi = (t_i4 - t_i) / 3, ri = t_i - pi;
    printf("  derived per-lane costs (arbitrary units): FP64 product %.3f reduction %.3f | IMAD product %.3f reduction %.3f\n", pf, rf, pi, ri);
    return 0;
}
