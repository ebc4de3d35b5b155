// pipes.cu -- micro-benchmarks of the two integer pipes the pairing kernels live on (sm_100a):
//   fmaheavy (IMAD.WIDE.U32[.X]) and alu (IADD3[.X]); can warps of one scheduler overlap them? what does large
//   straight-line code cost?  Numbers decide the occupancy / code-size strategy of k_miller_fexp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box, prints a table.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// 4 chained IMAD.WIDE.U32.X (one carry chain over an 8-limb window)
#define MADROW(a, x0, x1, x2, x3, y)                                                                  \
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"              \
                 "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"             \
                 "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"           \
                 "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"                  \
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) \
                 : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y))
// 8 chained IADD3.X (one 256-bit add)
#define ADD8(a, b)                                                                                    \
    asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\t" \
                 "addc.cc.u32 %3, %3, %11;\n\taddc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\t" \
                 "addc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"                                  \
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) \
                 : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]))

struct Regs {
    uint32_t m[4][8];  // four IMAD accumulator windows
    uint32_t s[4][8];  // four adder chains
    uint32_t x[4], y;
};
__device__ __forceinline__ void init(Regs& r, uint32_t seed) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        r.x[j] = seed * 2654435761u + j;
#pragma unroll
        for (int i = 0; i < 8; i++) { r.m[j][i] = seed + i * 7 + j; r.s[j][i] = seed * 3 + i + j * 11; }
    }
    r.y = seed * 40503u + 12345u;
}
__device__ __forceinline__ uint32_t fold(const Regs& r) {
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= r.m[j][i] ^ r.s[j][i];
    return s;
}
// one "unit": NM rows of IMAD.WIDE (4 each) on rotating windows and NA 8-limb adds on rotating chains
template <int NM, int NA>
__device__ __forceinline__ void unit(Regs& r) {
#pragma unroll
    for (int i = 0; i < (NM > NA ? NM : NA); i++) {
        if (i < NM) MADROW(r.m[i & 3], r.x[0], r.x[1], r.x[2], r.x[3], r.y + (i & 3));
        if (i < NA) ADD8(r.s[i & 3], r.s[(i + 1) & 3]);
    }
}

// mode 0: every warp runs unit<NM,NA>;  mode 1: even warps unit<NM,0>, odd warps unit<0,NA>
template <int NM, int NA, int MODE>
__global__ void k_mix(uint32_t* out, long long* cyc, int iters) {
    Regs r;
    init(r, threadIdx.x + blockIdx.x * blockDim.x);
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    long long t0 = clock64();
    if (MODE == 0) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) unit<NM, NA>(r);
    } else if ((warp >> 2) & 1) {   // warps 4..7 (second warp of every scheduler) add, warps 0..3 multiply
#pragma unroll 1
        for (int i = 0; i < iters; i++) unit<0, NA>(r);
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) unit<NM, 0>(r);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    uint32_t s = fold(r);
    if (s == 0x12345678u) out[0] = s;
}

// straight-line code of COPIES units (each unit<8,8> = 32 IMAD.WIDE + 64 IADD3 = 96 instr = 1.5 KB), looped; warps are
// de-phased by a per-warp prologue so that they sit at different program counters.
template <int COPIES>
__global__ void k_code(uint32_t* out, long long* cyc, int iters, int dephase) {
    Regs r;
    init(r, threadIdx.x + blockIdx.x * blockDim.x);
    const int warp = threadIdx.x >> 5;
    if (dephase) {
        int pre = (warp * 37 + blockIdx.x * 11) % 61;
#pragma unroll 1
        for (int i = 0; i < pre; i++) unit<8, 8>(r);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < COPIES; c++) unit<8, 8>(r);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    uint32_t s = fold(r);
    if (s == 0x12345678u) out[0] = s;
}

static uint32_t* d_out;
static long long* d_cyc;
static int n_sm;

template <class K, class... A>
static double run(K kern, int warps_per_smsp, int blocks_per_sm, A... args) {
    int threads = 128 * warps_per_smsp;
    int grid = n_sm * blocks_per_sm;
    kern<<<grid, threads>>>(d_out, d_cyc, args...);
    cudaDeviceSynchronize();
    kern<<<grid, threads>>>(d_out, d_cyc, args...);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return -1; }
    int nw = grid * threads / 32;
    std::vector<long long> h(nw);
    cudaMemcpy(h.data(), d_cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    return (double)h[nw / 2];   // median warp cycles
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    n_sm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, n_sm, p.clockRate);
    CK(cudaMalloc(&d_out, 4096));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * n_sm * 64 * 8));
    const int it = 2000;
    printf("\n[A] same code in every warp; cycles per unit per warp (W = warps per scheduler, one block per SM)\n");
    printf("%-34s %8s %8s %8s %8s\n", "unit", "W=1", "W=2", "W=3", "W=4");
#define ROW(NM, NA)                                                                                       \
    {                                                                                                     \
        printf("%2d IMAD.WIDE + %2d IADD3            ", 4 * NM, 8 * NA);                                  \
        for (int w = 1; w <= 4; w++) printf(" %8.1f", run(k_mix<NM, NA, 0>, w, 1, it) / it);              \
        printf("\n");                                                                                     \
    }
    ROW(8, 0) ROW(0, 8) ROW(8, 2) ROW(8, 4) ROW(8, 8) ROW(8, 16) ROW(4, 16) ROW(2, 16)
    printf("\n[B] heterogeneous warps on one scheduler: warp A 32 IMAD.WIDE per unit, warp B NA x 8 IADD3 per unit (W=2)\n");
    printf("    median warp cycles per unit (alone: rows above)\n");
    printf("  B=16 IADD3 : %8.1f\n", run(k_mix<8, 2, 1>, 2, 1, it) / it);
    printf("  B=32 IADD3 : %8.1f\n", run(k_mix<8, 4, 1>, 2, 1, it) / it);
    printf("  B=64 IADD3 : %8.1f\n", run(k_mix<8, 8, 1>, 2, 1, it) / it);
    printf("  B=128 IADD3: %8.1f\n", run(k_mix<8, 16, 1>, 2, 1, it) / it);
    printf("\n[C] straight-line code size (unit = 32 IMAD.WIDE + 64 IADD3 = 96 instr = 1.5 KB); cycles per unit per warp\n");
    printf("%-12s %10s %10s %10s %10s %10s %10s\n", "code", "W=2", "W=2 deph", "W=3", "W=3 deph", "W=4", "W=4 deph");
#define CROW(C)                                                                                           \
    {                                                                                                     \
        printf("%5.1f KB    ", C * 1.5);                                                                  \
        int iters = 4096 / C;                                                                             \
        for (int w = 2; w <= 4; w++)                                                                      \
            for (int d = 0; d < 2; d++) printf(" %10.1f", run(k_code<C>, w, 1, iters, d) / (iters * C)); \
        printf("\n");                                                                                     \
    }
    CROW(2) CROW(8) CROW(16) CROW(24) CROW(32) CROW(48) CROW(64) CROW(96)
    return 0;
}
