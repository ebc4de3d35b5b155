// kernels.cu -- sm_100a kernels and the C ABI (include/bn_b200.h) of the batched BN254 engine.
//
// Kernel map:
//   k_fq_mul_chain    K1  Fq Montgomery multiply chain (BASELINE config 2); k_imad_peak calibrates the IMAD.WIDE roofline
//   k_g1_mul/k_g2_mul K3  batched scalar multiplication, one thread per point (reference's chain: Jacobian limbs match)
//   k_pair_lines_duo  K4a to_affine (one inversion per pair, batched per block) + the ate line schedule (88 lines, NAF
//                         walk of 6u+2), one LANE PAIR per pairing (one Fq2 component per lane, duo.cuh), streamed to HBM
//                         in consumption order
//                         (k_pair_lines: the one-thread-per-pairing form, kept for A/B via BN_B200_LINES=solo)
//   k_miller          K4b Miller accumulation, one 6-lane hexad per pairing (5 per warp, 3 blocks/SM); Fq12 state in
//                         registers, operands exchanged through shared-memory slots (LDS/STS by 32-bit shared address),
//                         line coefficients read from a TMA-fed ring; writes the unreduced Miller value to the output
//   k_fexp            K4c final exponentiation in place on that buffer (2 blocks/SM; idle values parked in shared memory)
//   k_fexp_gather         same, the epilogue stores each result into every peer GPU's gather buffer over NVLink (row e)
//   k_fexp_pow            same + fused Gt::pow (row f-1);  k_miller_fexp: the fused single-kernel form (BN_SPLIT_KERNELS=0)
//   k_gt_mul/k_gt_pow/k_gt_inv K5 batched Gt arithmetic on hexads
//   k_fr_op, k_g{1,2}_{normalize,check,encode,decode}, k_fr_{encode,decode}   rows f-4 / f-3 (wire.cuh)
// There is no CPU fallback anywhere in this file: without a device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bn_b200.h"
#include "pairing.cuh"
#include "wire.cuh"

using namespace bn;

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fp ld_fp(const uint32_t* p) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
// plain ld.global (coherent): for buffers the same kernel also writes in place (the read-only path of __ldg is undefined there)
__device__ __forceinline__ Fp ld_fp_rw(const uint32_t* p) {
    const uint4 a = reinterpret_cast<const uint4*>(p)[0];
    const uint4 b = reinterpret_cast<const uint4*>(p)[1];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fp(uint32_t* p, const Fp& a) {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ void st128(uint32_t* p, const uint32_t* v) { reinterpret_cast<uint4*>(p)[0] = make_uint4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ Fp2 ld_fp2(const uint32_t* p) { return Fp2{ld_fp(p), ld_fp(p + 8)}; }
__device__ __forceinline__ Fp2 ld_fp2_rw(const uint32_t* p) { return Fp2{ld_fp_rw(p), ld_fp_rw(p + 8)}; }
__device__ __forceinline__ void st_fp2(uint32_t* p, const Fp2& a) {
    st_fp(p, a.c0);
    st_fp(p + 8, a.c1);
}

#ifndef HEX_WARPS_PER_BLOCK
#define HEX_WARPS_PER_BLOCK 4
#endif
#define HEX_PER_WARP 5
#define HEX_PER_BLOCK (HEX_WARPS_PER_BLOCK * HEX_PER_WARP)
#ifndef HEX_BATCH_INV
#define HEX_BATCH_INV 1   // final-exponentiation Fq inversion: one Fermat chain per block (1) or per lane (0)
#endif

// Shared scratch of a hexad block: one Fq slot per hexad + prefix products for the batched inversion, and the
// operand-exchange area: per lane three 64-byte slots (a, xi*a / aux, b / aux), lane stride padded to 52 words so the
// six lanes of a hexad hit disjoint bank groups with 128-bit accesses.
#if BN_F52
// FP64 operand exchange (f52.cuh): a slot holds a D5x3 = three 5-double factors at 48-byte pitch (16-byte aligned: two
// LDS.128 + one LDS.64 each); the integer form of a value (64 bytes, cyclotomic squaring / inversion) uses the first 64
// bytes of the same slot.  108-word lane stride: 108 mod 32 = 12, so any 8 consecutive lanes hit 8 disjoint 16-byte bank
// groups.
#define HEX_SLOT_BYTES 144
#define HEX_LANE_STRIDE 108
#else
#define HEX_SLOT_BYTES 64
#define HEX_LANE_STRIDE 52
#endif
#ifndef BN_LINE_TMA
#define BN_LINE_TMA 1   // stream the line coefficients HBM -> shared memory with cp.async.bulk (TMA) one step ahead
#endif
#define HEX_LINE_BYTES (HEX_PER_WARP * BN_LINE_WORDS * 4)   // one Miller step of one warp: 5 x 320 B, contiguous in HBM
struct alignas(128) LineRing {
    uint32_t buf[2][HEX_PER_WARP * BN_LINE_WORDS];
    unsigned long long bar[2];
};
struct alignas(128) HexSmem {
    Fp val[HEX_PER_BLOCK];
    Fp pre[HEX_PER_BLOCK];
    alignas(16) uint32_t xch[HEX_WARPS_PER_BLOCK][32 * HEX_LANE_STRIDE];
    alignas(16) uint32_t kq[16 * BN_KQ_STRIDE];  // k*q, k = 0..15 (xi-multiplication reduction rows)
    alignas(16) uint32_t zero[HEX_SLOT_BYTES / 4];  // an all-zero operand: "absent" for a lane role, selected by address
};
// dynamic shared memory of the hexad kernels: HexSmem | line rings (Miller kernels) | parking area (final exponentiation)
#if BN_LINE_TMA
#define HEX_RING_BYTES (HEX_WARPS_PER_BLOCK * sizeof(LineRing))
#else
#define HEX_RING_BYTES 0
#endif

// 1/x for one x per "slot" of a thread block with ONE inversion (Montgomery's simultaneous inversion):
// prefix products, invert the total, peel back.  Collective over the whole block (two __syncthreads); the chain
// runs on one thread while the other warps of the block yield their issue slots to the SM's other blocks.
// wslot: slot this thread writes (-1: none); rslot: slot it reads back (-1: slot 0, value unused).
// A zero input (only possible for infinity pairs / padding lanes, whose result is discarded) is replaced by 1 so it
// cannot poison the other elements' product.
__device__ __noinline__ Fp block_batch_inv(const Fp& x, int wslot, int rslot, int nslots, Fp* val, Fp* pre) {
    if (wslot >= 0) val[wslot] = fp_is_zero(x) ? fq_one() : x;
    __syncthreads();
    if (threadIdx.x == 0) {
        Fp acc = val[0];
        pre[0] = acc;
        for (int i = 1; i < nslots; i++) {
            acc = fp_mul<MQ>(acc, val[i]);
            pre[i] = acc;
        }
        Fp t = fq_inv_euclid(acc);  // single thread: the divergent binary Euclid is ~3x shorter than Fermat here
        for (int i = nslots - 1; i > 0; i--) {
            Fp vi = val[i];
            val[i] = fp_mul<MQ>(t, pre[i - 1]);
            t = fp_mul<MQ>(t, vi);
        }
        val[0] = t;
    }
    __syncthreads();
    return val[rslot >= 0 ? rslot : 0];
}

// Shared-memory access by 32-bit shared-window address (LDS.128 / STS.128): no generic-pointer arithmetic or address
// translation in the hexad round loops.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t* v) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void lds128d(uint32_t addr, double* v) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void lds64d(uint32_t addr, double* v) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[0]) : "r"(addr)); }
__device__ __forceinline__ void sts128d(uint32_t addr, const double* v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ void sts64d(uint32_t addr, const double* v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v[0]) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t* v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ Fp2 lds_fp2(uint32_t addr) {
    Fp2 r;
    lds128(addr, r.c0.v);
    lds128(addr + 16, r.c0.v + 4);
    lds128(addr + 32, r.c1.v);
    lds128(addr + 48, r.c1.v + 4);
    return r;
}
__device__ __forceinline__ void sts_fp2(uint32_t addr, const Fp2& v) {
    sts128(addr, v.c0.v);
    sts128(addr + 16, v.c0.v + 4);
    sts128(addr + 32, v.c1.v);
    sts128(addr + 48, v.c1.v + 4);
}
struct KqRowLds {  // row qhat of the k*q table (fp2.cuh: fp_small_reduce9)
    uint32_t base;
    __device__ __forceinline__ void operator()(uint32_t qhat, uint32_t* kq) const {
        const uint32_t a = base + qhat * (BN_KQ_STRIDE * 4);
        lds128(a, kq);
        lds128(a + 16, kq + 4);
    }
};

struct DevCtx {
    int kk;
    int slot;          // hexad index inside the block, or -1 for the two spare lanes of a warp
    HexSmem* sm;
    uint32_t mine;     // shared address of this lane's exchange slots
    uint32_t hexbase;  // shared address of lane 0 of this hexad
    uint32_t kq;       // shared address of the k*q table
    uint32_t zero;     // shared address of 64 zero bytes
    uint32_t parkbase; // shared address of this lane's parking area (0: kernel has none); layout [slot][16-byte piece][lane]
    __device__ __forceinline__ int k() const { return kk; }
    __device__ __forceinline__ void park(int s, const Fp2& v) const {
        const uint32_t a = parkbase + s * (4 * 32 * 16);
        sts128(a, v.c0.v);
        sts128(a + 512, v.c0.v + 4);
        sts128(a + 1024, v.c1.v);
        sts128(a + 1536, v.c1.v + 4);
    }
    __device__ __forceinline__ Fp2 unpark(int s) const {
        const uint32_t a = parkbase + s * (4 * 32 * 16);
        Fp2 r;
        lds128(a, r.c0.v);
        lds128(a + 512, r.c0.v + 4);
        lds128(a + 1024, r.c1.v);
        lds128(a + 1536, r.c1.v + 4);
        return r;
    }
    __device__ __forceinline__ Fp2 mul_xi(const Fp2& a) const { return fp2_mul_xi_r(a, KqRowLds{kq}); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ void put(int s, const Fp2& v) const { sts_fp2(mine + s * HEX_SLOT_BYTES, v); }
    __device__ __forceinline__ Fp2 get(int src, int s) const { return lds_fp2(hexbase + src * (HEX_LANE_STRIDE * 4) + s * HEX_SLOT_BYTES); }
    __device__ __forceinline__ Fp2 get_or_zero(bool cond, int src, int s) const {
        return lds_fp2(cond ? hexbase + src * (HEX_LANE_STRIDE * 4) + s * HEX_SLOT_BYTES : zero);
    }
#if BN_F52
    typedef uint32_t Ref;  // shared address of a published D5x3
    __device__ __forceinline__ void put52(int s, const Fp2& v) const {
        const D5x3 t = f52_from_fp2(v);
        const uint32_t a = mine + s * HEX_SLOT_BYTES;
        sts128d(a, t.c0.l); sts128d(a + 16, t.c0.l + 2); sts64d(a + 32, t.c0.l + 4);
        sts128d(a + 48, t.c1.l); sts128d(a + 64, t.c1.l + 2); sts64d(a + 80, t.c1.l + 4);
        sts128d(a + 96, t.cs.l); sts128d(a + 112, t.cs.l + 2); sts64d(a + 128, t.cs.l + 4);
    }
    __device__ __forceinline__ Ref ref(int src, int s) const { return hexbase + src * (HEX_LANE_STRIDE * 4) + s * HEX_SLOT_BYTES; }
    __device__ __forceinline__ Ref ref_or_zero(bool cond, int src, int s) const { return cond ? ref(src, s) : zero; }
    __device__ __forceinline__ D5 ld5(Ref r, int which) const {
        D5 v;
        const uint32_t a = r + which * 48;
        lds128d(a, v.l); lds128d(a + 16, v.l + 2); lds64d(a + 32, v.l + 4);
        return v;
    }
#endif
    __device__ __forceinline__ Fp small_reduce(const Lazy9& x) const { return lazy_reduce(x, KqRowLds{kq}); }
    __device__ __forceinline__ ModRegs mod_q() const {  // row 1 of the k*q table
        ModRegs r;
        KqRowLds{kq}(1u, r.p);
        return r;
    }
    __device__ __forceinline__ Fp inv(const Fp& x) const {
#if HEX_BATCH_INV
        return block_batch_inv(x, kk == 0 ? slot : -1, slot, HEX_PER_BLOCK, sm->val, sm->pre);
#else
        return fp_inv<MQ>(x);
#endif
    }
};

// lines live in HBM as [line t][pairing p][80 words]: a warp's five pairings read 5*320 contiguous bytes
struct DevLineSrc {
    const uint32_t* base;
    size_t n, pidx;
    int k;
    typedef const uint32_t* Handle;
    __device__ __forceinline__ Handle acquire(int t) const { return base + ((size_t)t * n + pidx) * BN_LINE_WORDS; }
    __device__ __forceinline__ Fp2 coef(Handle L, int i) const {
        const int off = i == 0 ? BN_LINE_OFF_L0 : i == 1 ? (k < 3 ? BN_LINE_OFF_XL3 : BN_LINE_OFF_L3) : (k < 4 ? BN_LINE_OFF_XL4 : BN_LINE_OFF_L4);
        return ld_fp2(L + off);
    }
    __device__ __forceinline__ void release(int) const {}
};
#if BN_LINE_TMA
// Line source fed by the TMA engine: one elected lane per warp issues a 1600-byte cp.async.bulk for Miller step t+2
// as soon as the warp has consumed step t; completion is tracked by an mbarrier transaction count, consumers spin on
// mbarrier.try_wait.parity.  A wait that exceeds its polling budget sets bit 0 of the library's device error word (the host
// turns it into BN_B200_ECUDA after the call) and the warp carries on with whatever the buffer holds: a protocol bug or a
// stalled copy engine fails THAT call with an error code; it neither hangs the GPU nor poisons the CUDA context (a __trap
// would).  BN_RING_TRAP=1 restores the trap for debugging.
struct DevLineSrcTma {
    const uint32_t* gbase;  // lines + p0 * 80 words (row 0 of this warp's five pairings)
    size_t row_words;       // n * 80
    uint32_t ring;          // shared address of this warp's LineRing (buf[0], buf[1], bar[0], bar[1])
    uint32_t lane_off;      // byte offset of this lane's hexad inside a ring buffer
    uint32_t coef_off;      // see coef()
    int lane;
    uint32_t* err;          // device error word (bit 0: line-ring wait timed out)
    __device__ __forceinline__ uint32_t bar(int b) const { return ring + 2 * HEX_LINE_BYTES + 8 * b; }
    __device__ __forceinline__ uint32_t buf(int b) const { return ring + b * HEX_LINE_BYTES; }
    __device__ __forceinline__ void init() const {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar(0)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar(1)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        issue(0);
        issue(1);
    }
    __device__ __forceinline__ void issue(int t) const {
        if (lane == 0) {
            const uint32_t* src = gbase + (size_t)t * row_words;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar(t & 1)), "r"((uint32_t)HEX_LINE_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf(t & 1)),
                         "l"(src), "r"((uint32_t)HEX_LINE_BYTES), "r"(bar(t & 1))
                         : "memory");
        }
    }
    typedef uint32_t Handle;  // shared address of this lane's line inside the ring buffer
    __device__ __forceinline__ Handle acquire(int t) const {
        const uint32_t parity = (uint32_t)(t >> 1) & 1u;
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 24) && !ok; spin++) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"(bar(t & 1)), "r"(parity)
                         : "memory");
            if (spin > (1 << 20)) __nanosleep(128);  // long wait: back off (about 2 s in total before giving up)
        }
        if (!ok) {
#if defined(BN_RING_TRAP) && BN_RING_TRAP
            __trap();
#else
            if (lane == 0) atomicOr(err, 1u);
#endif
        }
        return buf(t & 1) + lane_off;
    }
    // coef_off: byte offsets of l0 | l3k | l4k for this lane, packed 10 bits each
    __device__ __forceinline__ Fp2 coef(Handle L, int i) const { return lds_fp2(L + ((coef_off >> (10 * i)) & 0x3ffu)); }
    __device__ __forceinline__ void release(int t) const {
        __syncwarp();  // every lane has read buffer t&1 -> it can be refilled
        if (t + 2 < BN_NUM_LINES) issue(t + 2);
    }
};
#endif

struct DevLineSink {
    uint32_t* base;
    size_t n, pidx;
    __device__ __forceinline__ void operator()(int t, const Line& L) const {
        uint32_t* p = base + ((size_t)t * n + pidx) * BN_LINE_WORDS;
        st_fp2(p + BN_LINE_OFF_L0, L.l0);
        st_fp2(p + BN_LINE_OFF_L3, L.l3);
        st_fp2(p + BN_LINE_OFF_XL3, L.xl3);
        st_fp2(p + BN_LINE_OFF_L4, L.l4);
        st_fp2(p + BN_LINE_OFF_XL4, L.xl4);
    }
};

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fq_mul_chain(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                      uint32_t* __restrict__ out, size_t n, uint32_t iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = ld_fp(a + i * 8), y = ld_fp(b + i * 8);
#pragma unroll 2
    for (uint32_t k = 0; k < iters; k++) x = fp_mul<ModQ>(x, y);
    st_fp(out + i * 8, x);
}

// Dedicated Montgomery squaring chain x <- x^2 (108 IMAD.WIDE against 136 for a general product; fp_sqr in fp.cuh).
__global__ void __launch_bounds__(256) k_fq_sqr_chain(const uint32_t* __restrict__ a, uint32_t* __restrict__ out, size_t n, uint32_t iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = ld_fp(a + i * 8);
#pragma unroll 2
    for (uint32_t k = 0; k < iters; k++) x = fp_sqr<ModQ>(x);
    st_fp(out + i * 8, x);
}

// Calibration: nothing but independent IMAD.WIDE.U32 chains (8 accumulator pairs per thread), to measure the
// fma-pipe integer-multiply issue rate the pairing kernels are bounded by.  32 IMAD.WIDE per loop trip.
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* __restrict__ out, uint32_t iters, uint32_t seed) {
    // four independent accumulator windows, each fed by 4-IMAD carry chains (the exact instruction the field
    // arithmetic is made of: IMAD.WIDE.U32.X); 32 IMAD.WIDE per loop trip, no other arithmetic.
    uint32_t acc[4][8];
    uint32_t x[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        x[j] = threadIdx.x * 2654435761u + seed + j;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[j][i] = blockIdx.x + i * 7 + j;
    }
    uint32_t y = blockIdx.x * 40503u + 12345u;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int j = 0; j < 4; j++) mad_row4_nc(acc[j], x[0], x[1], x[2], x[3], y + j);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= acc[j][i];
    if (s == 0x12345678u && y == 3) out[0] = s;  // practically never: keeps the chains live
}

__global__ void __launch_bounds__(128) k_g1_mul(const uint32_t* __restrict__ p, const uint32_t* __restrict__ k,
                                                uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac<FqOps> P;
    P.x = ld_fp(p + i * 24);
    P.y = ld_fp(p + i * 24 + 8);
    P.z = ld_fp(p + i * 24 + 16);
    Jac<FqOps> r = jac_mul<FqOps>(P, ld_fp(k + i * 8));
    st_fp(out + i * 24, r.x);
    st_fp(out + i * 24 + 8, r.y);
    st_fp(out + i * 24 + 16, r.z);
}

__global__ void __launch_bounds__(128) k_g2_mul(const uint32_t* __restrict__ p, const uint32_t* __restrict__ k,
                                                uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac<Fq2Ops> P;
    P.x = ld_fp2(p + i * 48);
    P.y = ld_fp2(p + i * 48 + 16);
    P.z = ld_fp2(p + i * 48 + 32);
    Jac<Fq2Ops> r = jac_mul<Fq2Ops>(P, ld_fp(k + i * 8));
    st_fp2(out + i * 48, r.x);
    st_fp2(out + i * 48 + 16, r.y);
    st_fp2(out + i * 48 + 32, r.z);
}

// ---- row a11: the group law at the boundary (reference impl Add / Sub / Neg for G1, G2: src/lib.rs:97-114, 140-157 ->
// src/groups/mod.rs:272-347; double: :228-247; PartialEq: :83-109).  One thread per element; outputs are the same
// un-normalised Jacobian triples the crate produces.  op: 0 a + b, 1 a - b, 2 -a, 3 a.double()
template <class F>
struct JacIO;
template <>
struct JacIO<FqOps> {
    static constexpr int WORDS = 24;
    static __device__ __forceinline__ Jac<FqOps> ld(const uint32_t* p) { return Jac<FqOps>{ld_fp(p), ld_fp(p + 8), ld_fp(p + 16)}; }
    static __device__ __forceinline__ Jac<FqOps> ld_s(const uint32_t* p) { return Jac<FqOps>{ld_fp_rw(p), ld_fp_rw(p + 8), ld_fp_rw(p + 16)}; }
    static __device__ __forceinline__ void st(uint32_t* p, const Jac<FqOps>& a) { st_fp(p, a.x); st_fp(p + 8, a.y); st_fp(p + 16, a.z); }
};
template <>
struct JacIO<Fq2Ops> {
    static constexpr int WORDS = 48;
    static __device__ __forceinline__ Jac<Fq2Ops> ld(const uint32_t* p) { return Jac<Fq2Ops>{ld_fp2(p), ld_fp2(p + 16), ld_fp2(p + 32)}; }
    static __device__ __forceinline__ Jac<Fq2Ops> ld_s(const uint32_t* p) { return Jac<Fq2Ops>{ld_fp2_rw(p), ld_fp2_rw(p + 16), ld_fp2_rw(p + 32)}; }
    static __device__ __forceinline__ void st(uint32_t* p, const Jac<Fq2Ops>& a) { st_fp2(p, a.x); st_fp2(p + 16, a.y); st_fp2(p + 32, a.z); }
};
template <class F>
__device__ __forceinline__ void group_op(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef JacIO<F> IO;
    Jac<F> x = IO::ld(a + i * IO::WORDS), r;
    switch (op) {
        case 0: r = jac_add<F>(x, IO::ld(b + i * IO::WORDS)); break;
        case 1: r = jac_add<F>(x, jac_neg<F>(IO::ld(b + i * IO::WORDS))); break;
        case 2: r = jac_neg<F>(x); break;
        default: r = jac_double<F>(x); break;
    }
    IO::st(out + i * IO::WORDS, r);
}
__global__ void __launch_bounds__(128) k_g1_op(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n, int op) {
    group_op<FqOps>(a, b, out, n, op);
}
__global__ void __launch_bounds__(128) k_g2_op(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n, int op) {
    group_op<Fq2Ops>(a, b, out, n, op);
}
__global__ void __launch_bounds__(128) k_g1_eq(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = jac_eq<FqOps>(JacIO<FqOps>::ld(a + i * 24), JacIO<FqOps>::ld(b + i * 24)) ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_g2_eq(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = jac_eq<Fq2Ops>(JacIO<Fq2Ops>::ld(a + i * 48), JacIO<Fq2Ops>::ld(b + i * 48)) ? 1 : 0;
}

// ---- scalar multiplication with block-wide compaction of the additions (BASELINE config 3) ----------------------------
// The reference's chain (double every step, add on set bits; src/groups/mod.rs:250-270) is kept operation for operation so
// the Jacobian limbs match, but the ADDITIONS of one bit position are gathered over the whole thread block: the threads
// whose scalar has the bit set park their running point in shared memory, the first c threads of the block (c = number of
// set bits, about half) each perform one addition, and the owners read their result back.  Whole warps above c skip the
// 16-multiplication addition instead of executing it with half of their lanes masked off (one thread per point with
// per-lane branches costs 7 + 16 multiplications per step and warp; compacted, 7 + 16 * ceil(c / 32) / warps ~ 7 + 9).
#ifndef GMUL_THREADS_G1
#define GMUL_THREADS_G1 512   // run 14: 512 threads x 1 block 9.26 M/s; 256 x 2 8.97; 128 x 4 8.64; 256 x 1 (176 registers) 7.46; 768 x 1 (80 registers) 6.05
#endif
#ifndef GMUL_BLOCKS_G1
#define GMUL_BLOCKS_G1 1
#endif
#ifndef GMUL_THREADS_G2
#define GMUL_THREADS_G2 128
#endif
template <class F, int THREADS>
__device__ __forceinline__ void g_mul_compact(const uint32_t* __restrict__ p, const uint32_t* __restrict__ k, uint32_t* __restrict__ out,
                                              size_t n, uint32_t* smem) {
    typedef JacIO<F> IO;
    constexpr int WORDS = IO::WORDS;
    uint32_t* pbuf = smem;                                  // [THREADS][WORDS]: every thread's base point
    uint32_t* rbuf = smem + THREADS * WORDS;                // [THREADS][WORDS]: running points handed to the adders
    uint32_t* list = rbuf + THREADS * WORDS;                // [THREADS]: owners of the compacted additions
    uint32_t* wcount = list + THREADS;                      // two slot counters (double-buffered by step parity)
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t i = (size_t)blockIdx.x * THREADS + tid;
    const bool active = i < n;
    const size_t ii = active ? i : n - 1;
    IO::st(pbuf + tid * WORDS, IO::ld(p + ii * WORDS));
    const Fp sc = fp_from_mont<ModR>(ld_fp(k + ii * 8));  // U256::from(Fr), reference src/fields/fp.rs:15-22
    Jac<F> res;
    res.x = F::zero();
    res.y = F::one();
    res.z = F::zero();  // G::zero(), reference src/groups/mod.rs:208-214
    bool found_one = false;
    if (tid < 2) wcount[tid] = 0;
    __syncthreads();
    int step = 0;
    for (int w = 7; w >= 0; w--) {
        uint32_t bits = 0;
#pragma unroll
        for (int l = 0; l < 8; l++) bits = (w == l) ? sc.v[l] : bits;
        for (int b = 31; b >= 0; b--, step++) {
            if (found_one) res = jac_double<F>(res);
            const bool add = active && ((bits >> b) & 1u);
            const uint32_t m = __ballot_sync(0xffffffffu, add);
            // slot allocation: one shared-memory atomic per warp on a double-buffered counter (the order of the compacted
            // list is irrelevant), so a step needs two block barriers, not three
            uint32_t* counter = wcount + (step & 1);
            int base = 0;
            if (lane == 0 && m) base = (int)atomicAdd(counter, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (add) {
                list[base + __popc(m & ((1u << lane) - 1u))] = tid;
                IO::st(rbuf + tid * WORDS, res);
                found_one = true;
            }
            if (tid == 0) wcount[(step + 1) & 1] = 0;  // next step's counter (last read before the previous step's second barrier)
            __syncthreads();
            const int total = (int)*counter;
            if (tid < total) {  // whole warps beyond `total` skip the addition
                const int j = (int)list[tid];
                IO::st(rbuf + j * WORDS, jac_add<F>(IO::ld_s(rbuf + j * WORDS), IO::ld_s(pbuf + j * WORDS)));
            }
            __syncthreads();
            if (add) res = IO::ld_s(rbuf + tid * WORDS);
        }
    }
    if (active) IO::st(out + i * WORDS, res);
}
extern __shared__ __align__(16) uint32_t gmul_smem[];
#define GMUL_SMEM(F_WORDS, THREADS) ((size_t)(2 * (THREADS) * (F_WORDS) + (THREADS) + (THREADS) / 32) * 4)
__global__ void __launch_bounds__(GMUL_THREADS_G1, GMUL_BLOCKS_G1) k_g1_mul_c(const uint32_t* __restrict__ p, const uint32_t* __restrict__ k,
                                                              uint32_t* __restrict__ out, size_t n) {
    g_mul_compact<FqOps, GMUL_THREADS_G1>(p, k, out, n, gmul_smem);
}
__global__ void __launch_bounds__(GMUL_THREADS_G2) k_g2_mul_c(const uint32_t* __restrict__ p, const uint32_t* __restrict__ k,
                                                              uint32_t* __restrict__ out, size_t n) {
    g_mul_compact<Fq2Ops, GMUL_THREADS_G2>(p, k, out, n, gmul_smem);
}

// Fr::pow(self, exp: Fr) (reference src/lib.rs:24 -> FieldElement::pow, src/fields/mod.rs:35-46: the exponent is
// U256::from(exp), all 256 bits walked MSB first, squaring from the first iteration).
__global__ void __launch_bounds__(128) k_fr_pow(const uint32_t* __restrict__ a, const uint32_t* __restrict__ e, uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fp x = ld_fp(a + i * 8);
    const Fp k = fp_from_mont<ModR>(ld_fp(e + i * 8));
    Fp r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = FR_ONE_f(l);
    for (int w = 7; w >= 0; w--) {
        uint32_t bits = 0;
#pragma unroll
        for (int l = 0; l < 8; l++) bits = (w == l) ? k.v[l] : bits;
        for (int b = 31; b >= 0; b--) {
            r = fp_sqr<ModR>(r);
            if ((bits >> b) & 1u) r = fp_mul<ModR>(x, r);
        }
    }
    st_fp(out + i * 8, r);
}

// ---- rows f-3 / f-4: batched Fr arithmetic and Group::normalize ------------------------------------------------
// op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 inverse(a) (0 -> 0; the crate returns None, src/fields/fp.rs:103-112)
__global__ void __launch_bounds__(128) k_fr_op(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                               uint32_t* __restrict__ out, size_t n, int op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = ld_fp(a + i * 8), y = (op <= 2) ? ld_fp(b + i * 8) : fp_zero(), r;
    switch (op) {
        case 0: r = fp_mul<ModR>(x, y); break;
        case 1: r = fp_add<ModR>(x, y); break;
        case 2: r = fp_sub<ModR>(x, y); break;
        case 3: r = fp_neg<ModR>(x); break;
        default: r = fp_is_zero(x) ? x : fp_inv<ModR>(x); break;
    }
    st_fp(out + i * 8, r);
}
// Group::normalize (reference src/lib.rs:88-95, 131-138): affine coordinates with z = one; infinity is left as is.
__global__ void __launch_bounds__(128) k_g1_normalize(const uint32_t* __restrict__ p, uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = ld_fp(p + i * 24), y = ld_fp(p + i * 24 + 8), z = ld_fp(p + i * 24 + 16);
    if (!fp_is_zero(z)) {
        Fp zi = fp_inv<MQ>(z), zi2 = fp_mul<MQ>(zi, zi);
        x = fp_mul<MQ>(x, zi2);
        y = fp_mul<MQ>(y, fp_mul<MQ>(zi2, zi));
        z = fq_one();
    }
    st_fp(out + i * 24, x);
    st_fp(out + i * 24 + 8, y);
    st_fp(out + i * 24 + 16, z);
}
__global__ void __launch_bounds__(128) k_g2_normalize(const uint32_t* __restrict__ p, uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp2 x = ld_fp2(p + i * 48), y = ld_fp2(p + i * 48 + 16), z = ld_fp2(p + i * 48 + 32);
    if (!fp2_is_zero(z)) {
        Fp2 zi = fp2_inv(z), zi2 = fp2_sqr(zi);
        x = fp2_mul(x, zi2);
        y = fp2_mul(y, fp2_mul(zi2, zi));
        z = fp2_one();
    }
    st_fp2(out + i * 48, x);
    st_fp2(out + i * 48 + 16, y);
    st_fp2(out + i * 48 + 32, z);
}

// Decode-side validity checks (row f-3; reference AffineG::decode, src/groups/mod.rs:178-205): the point is given as
// (x, y, z = one) in Montgomery form (what to_jacobian() yields after Fq::new); ok = on the curve y^2 = x^3 + b and, for
// G2 (check_order() = true, :399), in the order-r subgroup: p * (-1) + p == zero.  z = 0 (the "00" encoding) is accepted.
__global__ void __launch_bounds__(128) k_g1_check(const uint32_t* __restrict__ p, uint8_t* __restrict__ ok, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = ld_fp(p + i * 24), y = ld_fp(p + i * 24 + 8), z = ld_fp(p + i * 24 + 16);
    Fp b;
#pragma unroll
    for (int l = 0; l < 8; l++) b.v[l] = G1_B_f(l);
    bool good = fp_eq(fp_mul<MQ>(y, y), fp_add<MQ>(fp_mul<MQ>(fp_mul<MQ>(x, x), x), b));
    ok[i] = (fp_is_zero(z) || good) ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_g2_check(const uint32_t* __restrict__ p, uint8_t* __restrict__ ok, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac<Fq2Ops> P;
    P.x = ld_fp2(p + i * 48);
    P.y = ld_fp2(p + i * 48 + 16);
    P.z = ld_fp2(p + i * 48 + 32);
    bool good = fp2_eq(fp2_sqr(P.y), fp2_add(fp2_mul(fp2_sqr(P.x), P.x), g2_coeff_b()));
    if (good && !fp2_is_zero(P.z)) {
        Fp m1;
#pragma unroll
        for (int l = 0; l < 8; l++) m1.v[l] = FR_MINUS_ONE_f(l);
        Jac<Fq2Ops> r = jac_add<Fq2Ops>(jac_mul<Fq2Ops>(P, m1), P);
        good = fp2_is_zero(r.z);
    }
    ok[i] = (fp2_is_zero(P.z) || good) ? 1 : 0;
}

// ---- row f-3: wire format (wire.cuh), one thread per element ------------------------------------------------------
// encode = Group::normalize + from-Montgomery + big-endian bytes (reference src/groups/mod.rs:143-163);
// decode = bytes -> Montgomery + the reference's validity checks, status per element (src/groups/mod.rs:178-205).
__global__ void __launch_bounds__(128) k_g1_encode(const uint32_t* __restrict__ p, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x = ld_fp(p + i * 24), y = ld_fp(p + i * 24 + 8), z = ld_fp(p + i * 24 + 16);
    const bool inf = fp_is_zero(z);
    if (!inf && !fp_eq(z, fq_one())) {
        Fp zi = fp_inv<MQ>(z), zi2 = fp_mul<MQ>(zi, zi);
        x = fp_mul<MQ>(x, zi2);
        y = fp_mul<MQ>(y, fp_mul<MQ>(zi2, zi));
    }
    g1_encode_affine(x, y, inf, out + i * 65);
}
__global__ void __launch_bounds__(128) k_g2_encode(const uint32_t* __restrict__ p, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp2 x = ld_fp2(p + i * 48), y = ld_fp2(p + i * 48 + 16), z = ld_fp2(p + i * 48 + 32);
    const bool inf = fp2_is_zero(z);
    if (!inf && !fp2_eq(z, fp2_one())) {
        Fp2 zi = fp2_inv(z), zi2 = fp2_sqr(zi);
        x = fp2_mul(x, zi2);
        y = fp2_mul(y, fp2_mul(zi2, zi));
    }
    g2_encode_affine(x, y, inf, out + i * 129);
}
__global__ void __launch_bounds__(128) k_g1_decode(const uint8_t* __restrict__ in, uint32_t* __restrict__ out, uint8_t* __restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac<FqOps> P;
    status[i] = g1_decode(in + i * 65, P);
    st_fp(out + i * 24, P.x);
    st_fp(out + i * 24 + 8, P.y);
    st_fp(out + i * 24 + 16, P.z);
}
__global__ void __launch_bounds__(128) k_g2_decode(const uint8_t* __restrict__ in, uint32_t* __restrict__ out, uint8_t* __restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac<Fq2Ops> P;
    status[i] = g2_decode(in + i * 129, P);
    st_fp2(out + i * 48, P.x);
    st_fp2(out + i * 48 + 16, P.y);
    st_fp2(out + i * 48 + 32, P.z);
}
__global__ void __launch_bounds__(128) k_fr_encode(const uint32_t* __restrict__ a, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_encode<ModR>(ld_fp(a + i * 8), out + i * 32);
}
__global__ void __launch_bounds__(128) k_fr_decode(const uint8_t* __restrict__ in, uint32_t* __restrict__ out, uint8_t* __restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x;
    const bool ok = fp_decode<ModR>(in + i * 32, x);
    status[i] = ok ? WIRE_OK : WIRE_NOT_REDUCED;
    st_fp(out + i * 8, ok ? x : fp_zero());
}

// K4a: one thread per pairing.  flags[p] = 1 when the pair is finite, 0 when either point is infinity.
__global__ void __launch_bounds__(64) k_pair_lines(const uint32_t* __restrict__ g1, const uint32_t* __restrict__ g2,
                                                   uint32_t* __restrict__ lines, uint8_t* __restrict__ flags, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Jac<FqOps> P;
    P.x = ld_fp(g1 + i * 24);
    P.y = ld_fp(g1 + i * 24 + 8);
    P.z = ld_fp(g1 + i * 24 + 16);
    Jac<Fq2Ops> Q;
    Q.x = ld_fp2(g2 + i * 48);
    Q.y = ld_fp2(g2 + i * 48 + 16);
    Q.z = ld_fp2(g2 + i * 48 + 32);
    Fp px, py;
    Fp2 qx, qy;
    SoloX X_;
    bool finite = pair_to_affine(X_, FermatInv(), P, Q, px, py, qx, qy);
    flags[i] = finite ? 1 : 0;
    DevLineSink sink{lines, n, i};
    ate_lines(X_, px, py, qx, qy, sink);
}

// K4a, lane-pair form (duo.cuh): lanes (2j, 2j+1) hold component 0 / 1 of every Fq2 value of pairing j.
// A pair's exchange area: 2 x 64 B (two operands per lane) + 16 B of padding that puts adjacent pairs on different banks.
// Every exchange is __syncwarp | write | __syncwarp | read: the first barrier says the partner has read the previous one.
#define DUO_XCH_WORDS 36
struct DevDuo {
    int hh;
    uint32_t kq;   // shared address of the k*q table
    uint32_t xch;  // shared address of this pair's exchange area
    __device__ __forceinline__ int h() const { return hh; }
    __device__ __forceinline__ void small_reduce9(uint32_t* v, uint32_t* out) const { fp_small_reduce9(v, out, KqRowLds{kq}); }
    __device__ __forceinline__ static void put(uint32_t addr, const Fp& v) {
        sts128(addr, v.v);
        sts128(addr + 16, v.v + 4);
    }
    __device__ __forceinline__ static Fp get(uint32_t addr) {
        Fp r;
        lds128(addr, r.v);
        lds128(addr + 16, r.v + 4);
        return r;
    }
    __device__ __forceinline__ void partner2(const Fp& a, const Fp& b, Fp& ao, Fp& bo) const {
        const uint32_t me = xch + (uint32_t)hh * 64, other = xch + (uint32_t)(hh ^ 1) * 64;
        __syncwarp();
        put(me, a);
        put(me + 32, b);
        __syncwarp();
        ao = get(other);
        bo = get(other + 32);
    }
    __device__ __forceinline__ Fp partner(const Fp& a) const {
        __syncwarp();
        put(xch + (uint32_t)hh * 64, a);
        __syncwarp();
        return get(xch + (uint32_t)(hh ^ 1) * 64);
    }
};
struct DevDuoLineSink {
    uint32_t* base;
    size_t n, pidx;
    int h;
    bool active;
    // lane h stores component h of the five coefficients
    __device__ __forceinline__ void operator()(int t, const LineH& L) const {
        if (!active) return;
        uint32_t* p = base + ((size_t)t * n + pidx) * BN_LINE_WORDS + 8 * h;
        st_fp(p + BN_LINE_OFF_L0, L.l0);
        st_fp(p + BN_LINE_OFF_L3, L.l3);
        st_fp(p + BN_LINE_OFF_XL3, L.xl3);
        st_fp(p + BN_LINE_OFF_L4, L.l4);
        st_fp(p + BN_LINE_OFF_XL4, L.xl4);
    }
};
#ifndef DUO_BLOCK
#define DUO_BLOCK 32   // threads per block (run 23: 32 -> 1.589 ms, 64 -> 1.613 ms, 128 -> 1.614 ms); DUO_BLOCK/2 pairings share one batched inversion
#endif
__global__ void __launch_bounds__(DUO_BLOCK) k_pair_lines_duo(const uint32_t* __restrict__ g1, const uint32_t* __restrict__ g2,
                                                       uint32_t* __restrict__ lines, uint8_t* __restrict__ flags, size_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i = t >> 1;
    const bool active = i < n;
    if (!active) i = n - 1;  // keep whole warps alive for the exchanges
    Jac<FqOps> P;
    P.x = ld_fp(g1 + i * 24);
    P.y = ld_fp(g1 + i * 24 + 8);
    P.z = ld_fp(g1 + i * 24 + 16);
    Jac<Fq2Ops> Q;
    Q.x = ld_fp2(g2 + i * 48);
    Q.y = ld_fp2(g2 + i * 48 + 16);
    Q.z = ld_fp2(g2 + i * 48 + 32);
    __shared__ alignas(16) uint32_t s_kq[16 * BN_KQ_STRIDE];  // k*q rows for the xi-multiplication's reduction
    __shared__ alignas(16) uint32_t s_xch[(DUO_BLOCK / 2) * DUO_XCH_WORDS];
    if (threadIdx.x < 16) kq_table_fill(s_kq, threadIdx.x);
    __syncthreads();
    const int slot = (int)(threadIdx.x >> 1);  // one slot per pairing of the block
    const int h = (int)(threadIdx.x & 1);
    DuoX<DevDuo> X_{DevDuo{h, smem_u32(s_kq), smem_u32(s_xch + slot * DUO_XCH_WORDS)}};
    Fp px, py;
    Fp2 qx, qy;
    __shared__ Fp s_val[DUO_BLOCK / 2], s_pre[DUO_BLOCK / 2];
    auto inv = [&](const Fp& x) { return block_batch_inv(x, h == 0 ? slot : -1, slot, DUO_BLOCK / 2, s_val, s_pre); };
    bool finite = pair_to_affine(X_, inv, P, Q, px, py, qx, qy);
    if (active && h == 0) flags[i] = finite ? 1 : 0;
    DevDuoLineSink sink{lines, n, i, h, active};
    ate_lines_duo(X_.d, px, py, qx, qy, sink);
}

#ifndef HEX_MIN_BLOCKS
#define HEX_MIN_BLOCKS 3   // blocks/SM promised to ptxas for k_miller_fexp (register cap = 65536 / (threads * blocks))
#endif
#ifndef HEX_MIN_BLOCKS_POW
#define HEX_MIN_BLOCKS_POW 1   // the fused pairing.pow kernel keeps four more Gt values live
#endif
#define HEX_PARK_WARP_BYTES (HX_PARK_SLOTS * 4 * 32 * 16)
#define HEX_PARK_BYTES (HEX_WARPS_PER_BLOCK * HEX_PARK_WARP_BYTES)
#define HEX_SMEM_PLAIN (sizeof(HexSmem))                                   // k_gt_*
#define HEX_SMEM_MILLER (sizeof(HexSmem) + HEX_RING_BYTES)                 // k_miller
#define HEX_SMEM_FEXP (sizeof(HexSmem) + HEX_PARK_BYTES)                   // k_fexp, k_fexp_pow, k_fexp_gather
#define HEX_SMEM_FUSED (sizeof(HexSmem) + HEX_RING_BYTES + HEX_PARK_BYTES) // k_miller_fexp: HexSmem | rings | parking
extern __shared__ __align__(128) unsigned char hex_dyn_smem[];

struct HexIndex {
    DevCtx ctx;
    size_t pidx;   // pairing / element index (clamped to a valid one)
    bool active;   // this lane belongs to a real element
};
__device__ __forceinline__ HexIndex hex_index(size_t n, HexSmem* sm) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int hex = lane / 6;  // 0..5 (5 = the two spare lanes)
    if (threadIdx.x < 16) kq_table_fill(sm->kq, threadIdx.x);
    for (int z = threadIdx.x; z < HEX_SLOT_BYTES / 4; z += blockDim.x) sm->zero[z] = 0;
    __syncthreads();
    HexIndex h;
    h.ctx.kk = lane - hex * 6;
    h.ctx.slot = hex < HEX_PER_WARP ? warp * HEX_PER_WARP + hex : -1;
    h.ctx.sm = sm;
    h.ctx.mine = smem_u32(sm->xch[warp] + lane * HEX_LANE_STRIDE);
    // the two spare lanes (30, 31) write their own slots but read hexad 4's, so every read stays inside the warp's area
    h.ctx.hexbase = smem_u32(sm->xch[warp] + (hex < HEX_PER_WARP ? hex : HEX_PER_WARP - 1) * 6 * HEX_LANE_STRIDE);
    h.ctx.kq = smem_u32(sm->kq);
    h.ctx.zero = smem_u32(sm->zero);
    h.ctx.parkbase = 0;
    size_t idx = ((size_t)blockIdx.x * HEX_WARPS_PER_BLOCK + warp) * HEX_PER_WARP + hex;
    h.active = (hex < HEX_PER_WARP) && (idx < n);
    h.pidx = h.active ? idx : (n - 1);
    return h;
}

// K4b: Miller loop and final exponentiation, one hexad per pairing, as TWO kernels (k_miller writes the unreduced
// Miller value into the output buffer, k_fexp finishes it in place): each kernel's hot code (about 30 KB) stays inside
// the SM's instruction cache and all resident warps run the same phase -- with the fused form (BN_SPLIT_KERNELS=0,
// k_miller_fexp) a third block per SM stalls on instruction fetch (profiles/README.md, runs 27-28).
// POW: additionally raise the result to the per-pairing scalar k (fused pairing(p, q).pow(k), row f-1).
#ifndef BN_SPLIT_KERNELS
#define BN_SPLIT_KERNELS 1
#endif
#ifndef MILLER_MIN_BLOCKS
#define MILLER_MIN_BLOCKS 2   // FP64 product path (round 2, run 6): 2.07 ms at 2 blocks/SM (214 registers) vs 2.15 ms at 3 (168); integer path (run 28): 3 was better
#endif
#ifndef FEXP_MIN_BLOCKS
#define FEXP_MIN_BLOCKS 2   // run 28: k_fexp 4.11 ms at 2 blocks/SM (254 registers) vs 4.37 ms at 3 (168 registers, L0 instruction-cache misses)
#endif
// dynamic shared memory: HexSmem, then (at park_off) the parking area (HX_PARK_SLOTS x 64 B per lane, hexad.cuh)
__device__ __forceinline__ HexIndex hex_index_dyn(size_t n, uint32_t park_off = sizeof(HexSmem)) {
    HexSmem* smem = reinterpret_cast<HexSmem*>(hex_dyn_smem);
    HexIndex h = hex_index(n, smem);
    h.ctx.parkbase = smem_u32(hex_dyn_smem + park_off) + (threadIdx.x >> 5) * HEX_PARK_WARP_BYTES + (threadIdx.x & 31) * 16;
    return h;
}
__device__ __forceinline__ Fp2 miller_part(const HexIndex& h, const uint32_t* __restrict__ lines, size_t n, uint32_t* err) {
#if BN_LINE_TMA
    LineRing* rings = reinterpret_cast<LineRing*>(hex_dyn_smem + sizeof(HexSmem));  // right after HexSmem (128-byte aligned)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    size_t p0 = ((size_t)blockIdx.x * HEX_WARPS_PER_BLOCK + warp) * HEX_PER_WARP;
    if (p0 >= n) p0 = n >= HEX_PER_WARP ? n - HEX_PER_WARP : 0;  // warp with no real pairing: any valid rows will do
    const int hex = lane / 6;
    DevLineSrcTma src{lines + p0 * BN_LINE_WORDS, n * (size_t)BN_LINE_WORDS, smem_u32(&rings[warp]),
                      (uint32_t)((hex < HEX_PER_WARP ? hex : HEX_PER_WARP - 1) * BN_LINE_WORDS * 4),
                      (uint32_t)(4 * BN_LINE_OFF_L0) | ((uint32_t)(4 * (h.ctx.kk < 3 ? BN_LINE_OFF_XL3 : BN_LINE_OFF_L3)) << 10) |
                          ((uint32_t)(4 * (h.ctx.kk < 4 ? BN_LINE_OFF_XL4 : BN_LINE_OFF_L4)) << 20),
                      lane, err};
    src.init();
#else
    DevLineSrc src{lines, n, h.pidx, h.ctx.kk};
#endif
    return hx_miller_loop(h.ctx, src);
}
// INV_BATCHED: f^-1 is finished from what k_miller prepared (aux: u0, u1, u2) and k_fq_inv_batch inverted (ninv);
// otherwise the whole inversion happens here, the Fq inversion batched over the block (Ctx::inv).
template <bool POW, bool INV_BATCHED>
__device__ __forceinline__ Fp2 fexp_part(const HexIndex& h, Fp2 f, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ k,
                                         const uint32_t* __restrict__ aux, const uint32_t* __restrict__ ninv) {
    if (INV_BATCHED) {
        const uint32_t* a = aux + h.pidx * 48;
        const Fp2 finv = hx_inv_finish(h.ctx, f, ld_fp2(a), ld_fp2(a + 16), ld_fp2(a + 32), ld_fp(ninv + h.pidx * 8));
        f = hx_final_exp_with_inverse(h.ctx, f, finv);
    } else {
        f = hx_final_exp(h.ctx, f);
    }
    if (!flags[h.pidx]) f = hx_one(h.ctx);  // infinity => Gt::one(), reference src/groups/mod.rs:765-766
    if (POW) {
        Fp e = fp_from_mont<ModR>(ld_fp(k + h.pidx * 8));  // U256::from(Fr), reference src/fields/fp.rs:15-22
        f = hx_pow_cyc(h.ctx, f, e);
    }
    return f;
}
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK, MILLER_MIN_BLOCKS)
k_miller(const uint32_t* __restrict__ lines, uint32_t* __restrict__ out, size_t n, uint32_t* err, uint32_t* __restrict__ aux,
         uint32_t* __restrict__ norm) {
    HexIndex h = hex_index_dyn(n);
    Fp2 f = miller_part(h, lines, n, err);
    if (h.active) st_fp2(out + h.pidx * 96 + 16 * gt_slot(h.ctx.kk), f);
    // first half of the final exponentiation's Fq12 inversion (hexad.cuh hx_inv_prepare): everything down to the one Fq
    // element per pairing that must be inverted.  All six lanes hold the same values; lanes 0..2 store u_k, lane 3 the norm.
    const HxInvPrep p = hx_inv_prepare(h.ctx, f);
    if (h.active) {
        const int kk = h.ctx.kk;
        if (kk < 3) st_fp2(aux + h.pidx * 48 + kk * 16, kk == 0 ? p.u0 : (kk == 1 ? p.u1 : p.u2));
        if (kk == 3) st_fp(norm + h.pidx * 8, p.nn);
    }
}

// v[i] <- v[i]^-1 (Montgomery form, mod q) for n values, one value per lane: a warp multiplies its 32 values together with
// a log-step prefix and suffix product scan (10 Fq multiplications of latency instead of 93), ONE lane inverts the total
// (binary Euclid, fp2.cuh: data-dependent loops run on a single lane), and every lane recovers its own inverse as
// total^-1 * prefix[i-1] * suffix[i+1].  Zero (only the padding / infinity pairs, whose results are discarded) counts as 1.
// Replaces the per-block single-thread inversion inside k_fexp, during which the block's four warps sat at a barrier
// (10 % of that kernel, profiles/r02_run5): 16 384 inversions cost one ~25 k-instruction Euclid chain of latency, once.
__global__ void __launch_bounds__(128) k_fq_inv_batch(uint32_t* __restrict__ v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    Fp x = fq_one();
    if (i < n) {
        x = ld_fp_rw(v + i * 8);
        if (fp_is_zero(x)) x = fq_one();
    }
    auto shfl_fp = [](const Fp& a, int src) {
        Fp r;
#pragma unroll
        for (int l = 0; l < 8; l++) r.v[l] = __shfl_sync(0xffffffffu, a.v[l], src);
        return r;
    };
    // inclusive scans: pre[i] = x_0 ... x_i, suf[i] = x_i ... x_31
    Fp pre = x, suf = x;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const Fp a = shfl_fp(pre, lane - d < 0 ? lane : lane - d), b = shfl_fp(suf, lane + d > 31 ? lane : lane + d);
        const Fp pa = fp_mul_ni<MQ>(pre, a), sb = fp_mul_ni<MQ>(suf, b);
        if (lane - d >= 0) pre = pa;
        if (lane + d <= 31) suf = sb;
    }
    Fp tinv = fq_one();
    if (lane == 31) tinv = fq_inv_euclid(pre);
    tinv = shfl_fp(tinv, 31);
    const Fp left = shfl_fp(pre, lane == 0 ? 0 : lane - 1), right = shfl_fp(suf, lane == 31 ? 31 : lane + 1);
    Fp r = tinv;
    if (lane > 0) r = fp_mul_ni<MQ>(r, left);
    if (lane < 31) r = fp_mul_ni<MQ>(r, right);
    if (i < n) st_fp(v + i * 8, r);
}
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK, FEXP_MIN_BLOCKS)
k_fexp(const uint8_t* __restrict__ flags, uint32_t* out, size_t n, const uint32_t* __restrict__ aux, const uint32_t* __restrict__ ninv) {
    HexIndex h = hex_index_dyn(n);
    uint32_t* p = out + h.pidx * 96 + 16 * gt_slot(h.ctx.kk);
    Fp2 f = fexp_part<false, true>(h, ld_fp2_rw(p), flags, nullptr, aux, ninv);
    if (h.active) st_fp2(p, f);
}
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK, HEX_MIN_BLOCKS_POW)
k_fexp_pow(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ k, uint32_t* out, size_t n, const uint32_t* __restrict__ aux,
           const uint32_t* __restrict__ ninv) {
    HexIndex h = hex_index_dyn(n);
    uint32_t* p = out + h.pidx * 96 + 16 * gt_slot(h.ctx.kk);
    Fp2 f = fexp_part<true, true>(h, ld_fp2_rw(p), flags, k, aux, ninv);
    if (h.active) st_fp2(p, f);
}
// Final exponentiation fused with the multi-GPU gather: the epilogue stores each result straight
// into every peer's gather buffer over NVLink (peer-mapped memory), at this rank's slot base; no separate all-gather.
// `out` = this rank's own slot (holds the Miller values written by k_miller).
#define BN_MAX_PEERS 8
struct PeerOut {
    uint32_t* slot[BN_MAX_PEERS];  // slot[r] = peer r's gather buffer + this rank's offset
};
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK, FEXP_MIN_BLOCKS)
k_fexp_gather(const uint8_t* __restrict__ flags, const uint32_t* out, PeerOut peers, int world, size_t n, const uint32_t* __restrict__ aux,
              const uint32_t* __restrict__ ninv) {
    HexIndex h = hex_index_dyn(n);
    const size_t off = h.pidx * 96 + 16 * gt_slot(h.ctx.kk);  // `out` may alias peers.slot[rank]: coherent loads, no __restrict__
    Fp2 f = fexp_part<false, true>(h, ld_fp2_rw(out + off), flags, nullptr, aux, ninv);
    if (h.active)
        for (int r = 0; r < world; r++) st_fp2(peers.slot[r] + off, f);
}
// fused single-kernel form (A/B: BN_SPLIT_KERNELS=0)
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK, HEX_MIN_BLOCKS)
k_miller_fexp(const uint32_t* __restrict__ lines, const uint8_t* __restrict__ flags, uint32_t* __restrict__ out, size_t n, uint32_t* err) {
    HexIndex h = hex_index_dyn(n, sizeof(HexSmem) + HEX_RING_BYTES);
    Fp2 f = fexp_part<false, false>(h, miller_part(h, lines, n, err), flags, nullptr, nullptr, nullptr);
    if (h.active) st_fp2(out + h.pidx * 96 + 16 * gt_slot(h.ctx.kk), f);
}

__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK)
k_gt_mul(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n) {
    HexIndex h = hex_index(n, reinterpret_cast<HexSmem*>(hex_dyn_smem));
    const int slot = 16 * gt_slot(h.ctx.kk);
    Fp2 x = ld_fp2(a + h.pidx * 96 + slot), y = ld_fp2(b + h.pidx * 96 + slot);
    Fp2 r = hx_mul(h.ctx, x, y);
    if (h.active) st_fp2(out + h.pidx * 96 + slot, r);
}

// Gt::inverse (reference src/lib.rs:172 -> Fq12::inverse, src/fields/fq12.rs:284-292); b is unused.
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK)
k_gt_inv(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n) {
    HexIndex h = hex_index(n, reinterpret_cast<HexSmem*>(hex_dyn_smem));
    const int slot = 16 * gt_slot(h.ctx.kk);
    Fp2 x = ld_fp2(a + h.pidx * 96 + slot);
    Fp2 r = hx_inv(h.ctx, x);
    if (h.active) st_fp2(out + h.pidx * 96 + slot, r);
}

// Fq12::exp_by_neg_z as the reference evaluates it (src/fields/fq12.rs:97-101, 229-246): binary cyclotomic_pow(u) with the
// literal Granger-Scott squaring, then conjugation -- defined for ANY Fq12 input (a polynomial map), which is what the
// reference's test_cyclotomic_exp pins (src/fields/mod.rs:171-201, a non-cyclotomic input).  b is unused.
__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK)
k_gt_exp_neg_z(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n) {
    HexIndex h = hex_index(n, reinterpret_cast<HexSmem*>(hex_dyn_smem));
    const int slot = 16 * gt_slot(h.ctx.kk);
    Fp2 x = ld_fp2(a + h.pidx * 96 + slot);
    Fp2 r = hx_exp_by_neg_z_literal(h.ctx, x);
    if (h.active) st_fp2(out + h.pidx * 96 + slot, r);
}

__global__ void __launch_bounds__(32 * HEX_WARPS_PER_BLOCK)
k_gt_pow(const uint32_t* __restrict__ a, const uint32_t* __restrict__ k, uint32_t* __restrict__ out, size_t n) {
    HexIndex h = hex_index(n, reinterpret_cast<HexSmem*>(hex_dyn_smem));
    const int slot = 16 * gt_slot(h.ctx.kk);
    Fp2 x = ld_fp2(a + h.pidx * 96 + slot);
    Fp e = fp_from_mont<ModR>(ld_fp(k + h.pidx * 8));  // U256::from(Fr), reference src/fields/fp.rs:15-22
    Fp2 r = hx_pow(h.ctx, x, e);
    if (h.active) st_fp2(out + h.pidx * 96 + slot, r);
}

#include "capi.inc"
