// pipes2.cu -- issue cost of every instruction the field arithmetic can be built from, on B200 (sm_100a), with register
// operands and asm volatile bodies (nothing can be hoisted or dead-code eliminated; the SASS of each probe is checked in
// profiles/r02_*_ubench_pipes2.txt by the opcode histogram this program's companion script prints).
//
//   per probe: 8 independent dependency chains per thread, 32 instructions per loop trip, W warps per scheduler
//   (W = 1, 2, 3, 4, 6, 8: one block of 128 W threads per SM, at most 64 registers per thread: no spills).
//   The number printed is SCHEDULER cycles per warp instruction = median warp cycles / (32 * trips) / 1 ... divided by
//   nothing else: at W warps the scheduler's throughput cost is (that number) / W, printed in the second table.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipes2 pipes2.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

enum Kind {
    K_IMAD_WIDE = 0,      // mad.wide.u32            IMAD.WIDE.U32  (no carry)
    K_IMAD_WIDE_X2,       // mad.lo.cc + madc.hi.cc + madc.lo.cc + madc.hi  = IMAD.WIDE.U32 then IMAD.WIDE.U32.X (2-chain)
    K_IMAD_WIDE_X4,       // the 4-chain of fp.cuh's rows (1 plain + 3 .X)
    K_IMAD_LO,            // mad.lo.u32              IMAD (32-bit)
    K_IMAD_HI,            // mad.hi.u32              IMAD.HI.U32
    K_IADD3,              // add.u32 with three register inputs (IADD3)
    K_IADD3_X,            // add.cc / addc chains of 8 (IADD3 + IADD3.X, fp.cuh add8)
    K_LOP3,               // lop3.b32
    K_SHF,                // shf.r.wrap.b32
    K_SEL,                // selp.b32
    K_DFMA,               // fma.rz.f64
    K_DADD,               // add.rz.f64
    K_FFMA,               // fma.rn.f32
    K_MIX_WIDE_IADD3,     // 16 IMAD.WIDE + 16 IADD3 interleaved
    K_MIX_DFMA_IADD3,     // 16 DFMA + 16 IADD3 interleaved
    K_MIX_DFMA_WIDE,      // 16 DFMA + 16 IMAD.WIDE interleaved
    K_MIX_ALL3,           // 11 DFMA + 11 IMAD.WIDE + 10 IADD3
    K_COUNT
};
static const char* NAMES[K_COUNT] = {
    "IMAD.WIDE.U32 (plain)", "IMAD.WIDE.U32 + .X (2-chains)", "IMAD.WIDE.U32.X (4-chains, fp.cuh)", "IMAD (32-bit lo)", "IMAD.HI.U32",
    "IADD3 (3 registers)", "IADD3.X carry chains of 8", "LOP3", "SHF", "SEL", "DFMA.RZ", "DADD.RZ", "FFMA",
    "mix 16 IMAD.WIDE + 16 IADD3", "mix 16 DFMA + 16 IADD3", "mix 16 DFMA + 16 IMAD.WIDE", "mix 11 DFMA + 11 IMAD.WIDE + 10 IADD3"};

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k_probe(unsigned long long* out, long long* cyc, int iters, unsigned seed) {
    unsigned acc[16];  // eight 64-bit accumulators as register pairs (acc[2i], acc[2i+1])
    unsigned a[8], b[8];
    double d[8], x, y;
    float f[8], fx, fy;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc[2 * i] = (i + 1) * 0x9E3779B9u + threadIdx.x + seed;
        acc[2 * i + 1] = (i + 1) * 0x7F4A7C15u + seed;
        a[i] = seed * (2 * i + 3) + threadIdx.x;
        b[i] = a[i] * 40503u + 7u;
        d[i] = 1.0 + 1e-3 * (i + threadIdx.x % 7);
        f[i] = 1.0f + 1e-3f * i;
    }
    x = 1.0 + 1e-9 * seed;
    y = 1e-7 * seed;
    fx = (float)x;
    fy = (float)y;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned* w = &acc[2 * i];
                if (KIND == K_IMAD_WIDE) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(w[0]), "+r"(w[1]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_IMAD_LO) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_IMAD_HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_IADD3) asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_SEL) asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %1, %0, p; }" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                if (KIND == K_DFMA) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(x), "d"(y));
                if (KIND == K_DADD) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(y));
                if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fx), "f"(fy));
                if (KIND == K_MIX_WIDE_IADD3) {
                    if (i & 1) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(w[0]), "+r"(w[1]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                    else asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                }
                if (KIND == K_MIX_DFMA_IADD3) {
                    if (i & 1) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(x), "d"(y));
                    else asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                }
                if (KIND == K_MIX_DFMA_WIDE) {
                    if (i & 1) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(x), "d"(y));
                    else asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(w[0]), "+r"(w[1]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                }
                if (KIND == K_MIX_ALL3) {
                    const int sel = (8 * r + i) % 3;
                    if (sel == 0) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(x), "d"(y));
                    else if (sel == 1) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(w[0]), "+r"(w[1]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                    else asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(w[0]) : "r"(a[i]), "r"(b[(i + r) & 7]));
                }
            }
            if (KIND == K_IMAD_WIDE_X2) {  // 16 two-instruction chains = 32 IMAD.WIDE
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    unsigned* w = &acc[4 * (i >> 1)];  // 4 limbs
#pragma unroll
                    for (int c = 0; c < 2; c++)
                        asm volatile("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.u32 %3, %5, %6, %3;"
                                     : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3])
                                     : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(b[(i + r + c) & 7]));
                }
            }
            if (KIND == K_IMAD_WIDE_X4) {  // 8 four-instruction chains = 32 IMAD.WIDE
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    unsigned* w = &acc[8 * (i & 1)];  // 8 limbs
                    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\tmadc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                                 "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\tmadc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                                 : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7])
                                 : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(a[(i + 2) & 7]), "r"(a[(i + 3) & 7]), "r"(b[(i + r) & 7]));
                }
            }
            if (KIND == K_IADD3_X) {  // 4 chains of 8 = 32 IADD3(.X)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    unsigned* w = &acc[8 * (c & 1)];
                    asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                                 "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                                 : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7])
                                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[(4 + c) & 7]), "r"(b[5]), "r"(b[6]), "r"(b[(r + c) & 7]));
                }
            }
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[(size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        // only the state a probe uses stays live across the timed loop (64-register budget: no spills)
        const bool uses_d = KIND == K_DFMA || KIND == K_DADD || KIND == K_MIX_DFMA_IADD3 || KIND == K_MIX_DFMA_WIDE || KIND == K_MIX_ALL3;
        if (uses_d) s ^= (unsigned long long)__double_as_longlong(d[i]);
        if (KIND == K_FFMA) s ^= __float_as_uint(f[i]);
        if (KIND != K_DFMA && KIND != K_DADD && KIND != K_FFMA) s ^= acc[2 * i] ^ ((unsigned long long)acc[2 * i + 1] << 32);
    }
    if (s == 0x123456789ULL) out[0] = s;
}

static int n_sm;
static unsigned long long* d_out;
static long long* d_cyc;

template <int KIND>
static double probe(int w, int iters) {
    // w warps per scheduler = 4 w warps per SM: one block of 128 w threads (w <= 8) or two blocks of 1024 (w = 16)
    int threads = 128 * w, blocks = 1;
    if (w > 8) return -1.0;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe<KIND>, threads, 0);
    if (occ < blocks) return -1.0;  // would not be co-resident: the number would be meaningless
    const int grid = n_sm * blocks;
    for (int rep = 0; rep < 2; rep++) {
        k_probe<KIND><<<grid, threads>>>(d_out, d_cyc, iters, 12345u + rep);
        if (cudaDeviceSynchronize() != cudaSuccess) return -2.0;
    }
    const int nw = grid * threads / 32;
    std::vector<long long> h(nw);
    cudaMemcpy(h.data(), d_cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    return (double)h[nw / 2] / iters / 32.0;
}

template <int KIND>
static void row(double* res) {
    const int ws[6] = {1, 2, 3, 4, 6, 8};
    for (int i = 0; i < 6; i++) res[i] = probe<KIND>(ws[i], 2000);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    n_sm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, n_sm, p.clockRate);
    CK(cudaMalloc(&d_out, 4096));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * n_sm * 64 * 2));
    static double res[K_COUNT][6];
    row<K_IMAD_WIDE>(res[K_IMAD_WIDE]);
    row<K_IMAD_WIDE_X2>(res[K_IMAD_WIDE_X2]);
    row<K_IMAD_WIDE_X4>(res[K_IMAD_WIDE_X4]);
    row<K_IMAD_LO>(res[K_IMAD_LO]);
    row<K_IMAD_HI>(res[K_IMAD_HI]);
    row<K_IADD3>(res[K_IADD3]);
    row<K_IADD3_X>(res[K_IADD3_X]);
    row<K_LOP3>(res[K_LOP3]);
    row<K_SHF>(res[K_SHF]);
    row<K_SEL>(res[K_SEL]);
    row<K_DFMA>(res[K_DFMA]);
    row<K_DADD>(res[K_DADD]);
    row<K_FFMA>(res[K_FFMA]);
    row<K_MIX_WIDE_IADD3>(res[K_MIX_WIDE_IADD3]);
    row<K_MIX_DFMA_IADD3>(res[K_MIX_DFMA_IADD3]);
    row<K_MIX_DFMA_WIDE>(res[K_MIX_DFMA_WIDE]);
    row<K_MIX_ALL3>(res[K_MIX_ALL3]);
    const int ws[6] = {1, 2, 3, 4, 6, 8};
    printf("\n[1] warp cycles per instruction (what ONE warp sees), W warps per scheduler\n%-40s", "instruction");
    for (int i = 0; i < 6; i++) printf(" %7s%-2d", "W=", ws[i]);
    printf("\n");
    for (int k = 0; k < K_COUNT; k++) {
        printf("%-40s", NAMES[k]);
        for (int i = 0; i < 6; i++) printf(" %9.2f", res[k][i]);
        printf("\n");
    }
    printf("\n[2] scheduler cycles per warp instruction = [1] / W: the issue cost when enough warps are resident\n%-40s", "instruction");
    for (int i = 0; i < 6; i++) printf(" %7s%-2d", "W=", ws[i]);
    printf("\n");
    for (int k = 0; k < K_COUNT; k++) {
        printf("%-40s", NAMES[k]);
        for (int i = 0; i < 6; i++) printf(" %9.2f", res[k][i] < 0 ? res[k][i] : res[k][i] / ws[i]);
        printf("\n");
    }
    printf("\n(negative: the launch would not have been co-resident / failed.  The IADD3 probe is two adds per statement and the\n"
           " SEL probe a SETP + SEL pair: read those rows as cost per PAIR.)\n");
    return 0;
}
