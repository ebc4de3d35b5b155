// dfma.cu -- is FP64 FMA a cheaper multiplier than IMAD.WIDE on B200?  Measures issue cycles per warp instruction for
// DFMA chains, IMAD.WIDE chains, FFMA chains and mixes (W warps per scheduler, one block per SM), to size the
// "52-bit limbs on the FP64 pipe" alternative discussed in DESIGN.md (next steps).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

template <int ND, int NI, int NF, int NA = 0>
__global__ void k(double* out, long long* cyc, int iters, double seed) {
    double d[8];
    float f[8];
    unsigned long long m[8];
    unsigned s_[4][8];
    unsigned a = threadIdx.x * 2654435761u + 12345u, b = threadIdx.x * 40503u + 7u;
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = seed + i + threadIdx.x; f[i] = (float)(seed + i); m[i] = i + threadIdx.x; }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) s_[j][i] = a * (j + 3) + i;
    const double x = seed * 1.0000001, y = seed * 0.9999999;
    const float fx = (float)x, fy = (float)y;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < ND) d[i] = fma(d[i], x, y);
                if (i < NI) m[i] = (unsigned long long)(a + i) * (unsigned long long)(b + r) + m[i];
                if (i < NF) f[i] = fmaf(f[i], fx, fy);
            }
#pragma unroll
            for (int i = 0; i < NA; i++) {
                unsigned* p = s_[i & 3]; const unsigned* q = s_[(i + 1) & 3];
                asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                    "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                    : "+r"(p[0]), "+r"(p[1]), "+r"(p[2]), "+r"(p[3]), "+r"(p[4]), "+r"(p[5]), "+r"(p[6]), "+r"(p[7])
                    : "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]));
            }
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i] + (double)f[i] + (double)m[i] + (double)(s_[0][i] ^ s_[1][i] ^ s_[2][i] ^ s_[3][i]);
    if (s == 1.2345) out[0] = s;
}

static double* d_out; static long long* d_cyc; static int n_sm;
template <class K> static double run(K kern, int w, int iters) {
    int threads = 128 * w;
    kern<<<n_sm, threads>>>(d_out, d_cyc, iters, 1.5);
    cudaDeviceSynchronize();
    kern<<<n_sm, threads>>>(d_out, d_cyc, iters, 1.5);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return -1; }
    int nw = n_sm * threads / 32;
    std::vector<long long> h(nw);
    cudaMemcpy(h.data(), d_cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    return (double)h[nw / 2] / iters;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); n_sm = p.multiProcessorCount;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_cyc, sizeof(long long) * n_sm * 64);
    printf("cycles per loop trip per warp (trip = 4 x [ND DFMA + NI IMAD.WIDE + NF FFMA], 8 independent chains each)\n");
    printf("%-34s %8s %8s %8s %8s\n", "trip", "W=1", "W=2", "W=4", "W=8");
#define ROW(ND, NI, NF) { printf("%2d DFMA + %2d IMAD.WIDE + %2d FFMA      ", 4*ND, 4*NI, 4*NF); \
    for (int w : {1, 2, 4, 8}) printf(" %8.1f", run(k<ND, NI, NF>, w, 2000)); printf("\n"); }
#define ROWA(ND, NI, NA) { printf("%2d DFMA + %2d IMAD.WIDE + %2d IADD3     ", 4*ND, 4*NI, 32*NA); \
    for (int w : {1, 2, 4, 8}) printf(" %8.1f", run(k<ND, NI, 0, NA>, w, 2000)); printf("\n"); }
    ROWA(0, 0, 2) ROWA(8, 0, 2) ROWA(8, 0, 4) ROWA(8, 8, 2) ROWA(0, 8, 2) ROWA(8, 4, 4)
    ROW(8, 0, 0) ROW(0, 8, 0) ROW(0, 0, 8) ROW(8, 8, 0) ROW(8, 0, 8) ROW(0, 8, 8) ROW(8, 8, 8) ROW(4, 8, 0) ROW(8, 4, 0)
    return 0;
}
