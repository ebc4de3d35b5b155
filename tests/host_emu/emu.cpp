// Host emulation of the kernels' arithmetic (TEST INFRASTRUCTURE, never shipped, never a fallback).
//
// Compiles bn_b200/csrc/*.cuh with g++: the PTX carry chains are replaced by their portable C
// equivalents (fp.cuh, #else branches) and a hexad's six lanes are six host threads that exchange
// registers through a barrier instead of __shfl_sync.  tests/test_host_emu.py checks these entry points
// against the oracle, so the tower / lane choreography / line schedule is validated without a GPU;
// the -m gpu tests then validate the real kernels (PTX included) through the C ABI.
#include <atomic>
#include <cfenv>
#include <cstring>
#include <thread>
#include <vector>

#include "../../bn_b200/csrc/pairing.cuh"  // pulls in fp.cuh, fp2.cuh, duo.cuh, curve.cuh, hexad.cuh
#include "../../bn_b200/csrc/wire.cuh"

using namespace bn;

namespace {

struct Barrier {
    std::atomic<int> count{0};
    std::atomic<int> gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        int g = gen.load();
        if (count.fetch_add(1) + 1 == n) {
            count.store(0);
            gen.fetch_add(1);
        } else {
            while (gen.load() == g) std::this_thread::yield();
        }
    }
};

struct HostHexShared {
    Fp2 slot[6][3];
    D5x3 slot52[6][3];
    Fp2 parked[6][HX_PARK_SLOTS];
    Barrier bar{6};
};

struct HostCtx {
    int kk;
    HostHexShared* sh;
    int k() const { return kk; }
    Fp inv(const Fp& x) const { return fp_inv<ModQ>(x); }
    Fp2 mul_xi(const Fp2& a) const {  // exercise the table-driven reduction the kernels use
        static uint32_t tab[16 * BN_KQ_STRIDE];
        static bool init = [] { for (int k = 0; k < 16; k++) kq_table_fill(tab, k); return true; }();
        (void)init;
        return fp2_mul_xi_t(a, tab);
    }
    static const uint32_t* kq_tab() {
        static uint32_t tab[16 * BN_KQ_STRIDE];
        static bool init = [] { for (int k = 0; k < 16; k++) kq_table_fill(tab, k); return true; }();
        (void)init;
        return tab;
    }
    ModRegs mod_q() const { return mod_regs<MQ>(); }
    Fp small_reduce(const Lazy9& x) const { return lazy_reduce(x, KqRowPtr{kq_tab()}); }
    Fp2 get_or_zero(bool cond, int src, int s) const { return cond ? sh->slot[src][s] : fp2_zero(); }
    void put(int s, const Fp2& v) const { sh->slot[kk][s] = v; }
    // FP64 operand exchange (f52.cuh): Ref = pointer to a published D5x3 (nullptr: the zero operand)
    typedef const D5x3* Ref;
    void put52(int s, const Fp2& v) const { sh->slot52[kk][s] = f52_from_fp2(v); }
    Ref ref(int src, int s) const { return &sh->slot52[src][s]; }
    Ref ref_or_zero(bool cond, int src, int s) const { return cond ? &sh->slot52[src][s] : nullptr; }
    D5 ld5(Ref r, int which) const {
        if (!r) return D5{{0.0, 0.0, 0.0, 0.0, 0.0}};
        return which == 0 ? r->c0 : (which == 1 ? r->c1 : r->cs);
    }
    Fp2 get(int src, int s) const { return sh->slot[src][s]; }
    void sync() const { sh->bar.wait(); }
    void park(int s, const Fp2& v) const { sh->parked[kk][s] = v; }
    Fp2 unpark(int s) const { return sh->parked[kk][s]; }
};

struct HostDuoShared {
    Fp pre[2][2];
    Barrier bar{2};
};
struct HostDuo {  // duo.cuh's lane-pair context: lane h owns component h
    int hh;
    HostDuoShared* sh;
    int h() const { return hh; }
    void small_reduce9(uint32_t* v, uint32_t* out) const {
        static uint32_t tab[16 * BN_KQ_STRIDE];
        static bool init = [] { for (int k = 0; k < 16; k++) kq_table_fill(tab, k); return true; }();
        (void)init;
        fp_small_reduce9(v, out, KqRowPtr{tab});
    }
    void partner2(const Fp& a, const Fp& b, Fp& ao, Fp& bo) const {
        sh->pre[hh][0] = a;
        sh->pre[hh][1] = b;
        sh->bar.wait();
        ao = sh->pre[hh ^ 1][0];
        bo = sh->pre[hh ^ 1][1];
        sh->bar.wait();
    }
    Fp partner(const Fp& a) const {
        Fp ao, bo;
        partner2(a, a, ao, bo);
        return ao;
    }
};

Fp load_fp(const uint64_t* p) {
    Fp r;
    for (int i = 0; i < 4; i++) {
        r.v[2 * i] = (uint32_t)p[i];
        r.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
    }
    return r;
}
void store_fp(uint64_t* p, const Fp& a) {
    for (int i = 0; i < 4; i++) p[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
Fp2 load_fp2(const uint64_t* p) { return Fp2{load_fp(p), load_fp(p + 4)}; }
void store_fp2(uint64_t* p, const Fp2& a) {
    store_fp(p, a.c0);
    store_fp(p + 4, a.c1);
}

// run fn(ctx) on six lanes
template <class Fn>
void run_hexad(Fn fn) {
    HostHexShared sh;
    std::vector<std::thread> th;
    for (int k = 0; k < 6; k++)
        th.emplace_back([&sh, k, &fn]() {
            std::fesetround(FE_TOWARDZERO);  // f52.cuh's fma_rz on the host: the FPU rounding mode is per thread
            HostCtx c{k, &sh};
            fn(c);
        });
    for (auto& t : th) t.join();
}

struct HostLineSrc {
    const uint64_t* lines;  // [BN_NUM_LINES][40] u64
    int k;
    typedef const uint64_t* Handle;
    Handle acquire(int t) const { return lines + (size_t)t * 40; }
    Fp2 coef(Handle L, int i) const {
        if (i == 0) return load_fp2(L + BN_LINE_OFF_L0 / 2);
        if (i == 1) return load_fp2(L + (k < 3 ? BN_LINE_OFF_XL3 : BN_LINE_OFF_L3) / 2);
        return load_fp2(L + (k < 4 ? BN_LINE_OFF_XL4 : BN_LINE_OFF_L4) / 2);
    }
    void release(int) const {}
};

struct HostLineSink {
    uint64_t* out;
    void operator()(int t, const Line& L) {
        uint64_t* p = out + (size_t)t * 40;
        store_fp2(p + 0, L.l0);
        store_fp2(p + 8, L.l3);
        store_fp2(p + 16, L.xl3);
        store_fp2(p + 24, L.l4);
        store_fp2(p + 32, L.xl4);
    }
};

Jac<FqOps> load_g1(const uint64_t* p) { return Jac<FqOps>{load_fp(p), load_fp(p + 4), load_fp(p + 8)}; }
Jac<Fq2Ops> load_g2(const uint64_t* p) { return Jac<Fq2Ops>{load_fp2(p), load_fp2(p + 8), load_fp2(p + 16)}; }

}  // namespace

extern "C" {

int emu_num_lines() { return BN_NUM_LINES; }
int emu_ate_naf() { return BN_ATE_NAF; }

// op: 0 mul, 1 add, 2 sub, 3 neg(a), 4 inv(a), 5 half(a), 6 from_mont(a);  which: 0 Fq, 1 Fr
void emu_fp_op(int op, int which, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    Fp x = load_fp(a), y = b ? load_fp(b) : fp_zero(), r;
    if (which == 0) {
        switch (op) {
            case 0: r = fp_mul<ModQ>(x, y); break;
            case 1: r = fp_add<ModQ>(x, y); break;
            case 2: r = fp_sub<ModQ>(x, y); break;
            case 3: r = fp_neg<ModQ>(x); break;
            case 4: r = fp_inv<ModQ>(x); break;
            case 5: r = fp_half<ModQ>(x); break;
            case 7: r = fq_inv_euclid(x); break;
            case 8: r = fp_sqr<ModQ>(x); break;
            default: r = fp_from_mont<ModQ>(x); break;
        }
    } else {
        switch (op) {
            case 0: r = fp_mul<ModR>(x, y); break;
            case 1: r = fp_add<ModR>(x, y); break;
            case 2: r = fp_sub<ModR>(x, y); break;
            case 3: r = fp_neg<ModR>(x); break;
            case 4: r = fp_inv<ModR>(x); break;
            case 5: r = fp_half<ModR>(x); break;
            case 8: r = fp_sqr<ModR>(x); break;
            default: r = fp_from_mont<ModR>(x); break;
        }
    }
    store_fp(out, r);
}
// (a*b + c*d) * 2^-256 mod q with interleaved product / reduction rows (fp.cuh mul_reduce_rows); operands may be lazy
// (<= q, or < 2q for a single product) as in duo.cuh's callers
void emu_fp_mul2(const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, uint64_t* out) {
    store_fp(out, fp_mul2<ModQ>(load_fp(a), load_fp(b), load_fp(c), load_fp(d)));
}
// a * b * 2^-256 mod q through the FP64 path (f52.cuh), to be compared with emu_fp_op(0, 0, ...)
void emu_f52_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    const int old = std::fegetround();
    std::fesetround(FE_TOWARDZERO);
    store_fp(out, f52_fp_mul(load_fp(a), load_fp(b)));
    std::fesetround(old);
}
int emu_f52_enabled() { return BN_F52; }
// op: 0 mul, 1 sqr, 2 mul_xi, 3 inv, 4 add, 5 sub
void emu_fp2_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    Fp2 x = load_fp2(a), y = b ? load_fp2(b) : fp2_zero(), r;
    switch (op) {
        case 0: r = fp2_mul(x, y); break;
        case 1: r = fp2_sqr(x); break;
        case 2: r = fp2_mul_xi(x); break;
        case 3: r = fp2_inv(x); break;
        case 4: r = fp2_add(x, y); break;
        case 6: {
            static uint32_t tab[16 * BN_KQ_STRIDE];
            for (int k = 0; k < 16; k++) kq_table_fill(tab, k);
            r = fp2_mul_xi_t(x, tab);
            break;
        }
        default: r = fp2_sub(x, y); break;
    }
    store_fp2(out, r);
}
void emu_g1_mul(const uint64_t* p, const uint64_t* fr, uint64_t* out) {
    Jac<FqOps> r = jac_mul<FqOps>(load_g1(p), load_fp(fr));
    store_fp(out, r.x); store_fp(out + 4, r.y); store_fp(out + 8, r.z);
}
void emu_g2_mul(const uint64_t* p, const uint64_t* fr, uint64_t* out) {
    Jac<Fq2Ops> r = jac_mul<Fq2Ops>(load_g2(p), load_fp(fr));
    store_fp2(out, r.x); store_fp2(out + 8, r.y); store_fp2(out + 16, r.z);
}
// lines: [BN_NUM_LINES][40] u64.  returns 1 if finite, 0 if either input is infinity.
int emu_lines(const uint64_t* g1, const uint64_t* g2, uint64_t* lines, uint64_t* p_affine8, uint64_t* q_affine16) {
    Fp px, py; Fp2 qx, qy;
    SoloX X_;
    bool ok = pair_to_affine(X_, FermatInv(), load_g1(g1), load_g2(g2), px, py, qx, qy);
    if (p_affine8) { store_fp(p_affine8, px); store_fp(p_affine8 + 4, py); }
    if (q_affine16) { store_fp2(q_affine16, qx); store_fp2(q_affine16 + 8, qy); }
    HostLineSink sink{lines};
    ate_lines(X_, px, py, qx, qy, sink);
    return ok ? 1 : 0;
}
// same, computed by a lane PAIR (duo.cuh): two host threads; lane h stores component h of every coefficient, as the
// device lanes do
struct HostDuoLineSink {
    uint64_t* out;
    int h;
    void operator()(int t, const LineH& L) {
        uint64_t* p = out + (size_t)t * 40 + 4 * h;
        store_fp(p, L.l0);
        store_fp(p + 8, L.l3);
        store_fp(p + 16, L.xl3);
        store_fp(p + 24, L.l4);
        store_fp(p + 32, L.xl4);
    }
};
int emu_lines_duo(const uint64_t* g1, const uint64_t* g2, uint64_t* lines) {
    HostDuoShared sh;
    int ok[2] = {0, 0};
    std::vector<std::thread> th;
    for (int h = 0; h < 2; h++)
        th.emplace_back([&, h]() {
            DuoX<HostDuo> X_{HostDuo{h, &sh}};
            Fp px, py; Fp2 qx, qy;
            ok[h] = pair_to_affine(X_, FermatInv(), load_g1(g1), load_g2(g2), px, py, qx, qy) ? 1 : 0;
            HostDuoLineSink sink{lines, h};
            ate_lines_duo(X_.d, px, py, qx, qy, sink);
        });
    for (auto& t : th) t.join();
    return ok[0] & ok[1];
}
// Gt images are bn::Gt layout (48 u64).  op: 0 mul(a,b) 1 sqr 2 cyc_sqr 3 inv 4 frob(p=arg) 5 exp_by_neg_z
//   6 final_exp 7 conj 8 pow(a, plain exponent b[0..3]) 9 pow_cyc(a, plain exponent)
void emu_gt_op(int op, const uint64_t* a, const uint64_t* b, int arg, uint64_t* out) {
    run_hexad([&](HostCtx& c) {
        Fp2 x = load_fp2(a + 8 * gt_slot(c.k()));
        Fp2 y = b && op == 0 ? load_fp2(b + 8 * gt_slot(c.k())) : fp2_zero();
        Fp2 r;
        switch (op) {
            case 0: r = hx_mul(c, x, y); break;
            case 1: r = hx_sqr(c, x); break;
            case 2: r = hx_cyc_sqr(c, x); break;
            case 3: r = hx_inv(c, x); break;
            case 4: r = hx_frob(c, x, arg); break;
            case 5: r = hx_exp_by_neg_z(c, x); break;
            case 6: r = hx_final_exp(c, x); break;
            case 7: r = hx_conj(c, x); break;
            case 9: r = hx_pow_cyc(c, x, load_fp(b)); break;
            default: r = hx_pow(c, x, load_fp(b)); break;
        }
        store_fp2(out + 8 * gt_slot(c.k()), r);
    });
}
// Miller loop only (unreduced), from stored lines
void emu_miller(const uint64_t* lines, uint64_t* out) {
    run_hexad([&](HostCtx& c) {
        HostLineSrc src{lines, c.k()};
        Fp2 f = hx_miller_loop(c, src);
        store_fp2(out + 8 * gt_slot(c.k()), f);
    });
}
// full pairing through the same code path as the kernels
void emu_pairing(const uint64_t* g1, const uint64_t* g2, uint64_t* out) {
    std::vector<uint64_t> lines(BN_NUM_LINES * 40);
    int finite = emu_lines(g1, g2, lines.data(), nullptr, nullptr);
    run_hexad([&](HostCtx& c) {
        HostLineSrc src{lines.data(), c.k()};
        Fp2 f = hx_miller_loop(c, src);
        f = hx_final_exp(c, f);
        if (!finite) f = hx_one(c);
        store_fp2(out + 8 * gt_slot(c.k()), f);
    });
}
// wire format (wire.cuh).  kind: 0 Fr, 1 G1, 2 G2.  encode: image (u64 LE words) -> record; decode: record -> image,
// returns the status byte.
void emu_wire_encode(int kind, const uint64_t* img, uint8_t* rec) {
    if (kind == 0) {
        fp_encode<ModR>(load_fp(img), rec);
    } else if (kind == 1) {
        Jac<FqOps> p = load_g1(img);
        const bool inf = fp_is_zero(p.z);
        Fp x = p.x, y = p.y;
        if (!inf && !fp_eq(p.z, fq_one())) {
            Fp zi = fp_inv<MQ>(p.z), zi2 = fp_mul<MQ>(zi, zi);
            x = fp_mul<MQ>(x, zi2);
            y = fp_mul<MQ>(y, fp_mul<MQ>(zi2, zi));
        }
        g1_encode_affine(x, y, inf, rec);
    } else {
        Jac<Fq2Ops> p = load_g2(img);
        const bool inf = fp2_is_zero(p.z);
        Fp2 x = p.x, y = p.y;
        if (!inf && !fp2_eq(p.z, fp2_one())) {
            Fp2 zi = fp2_inv(p.z), zi2 = fp2_sqr(zi);
            x = fp2_mul(x, zi2);
            y = fp2_mul(y, fp2_mul(zi2, zi));
        }
        g2_encode_affine(x, y, inf, rec);
    }
}
int emu_wire_decode(int kind, const uint8_t* rec, uint64_t* img) {
    if (kind == 0) {
        Fp x;
        const bool ok = fp_decode<ModR>(rec, x);
        store_fp(img, ok ? x : fp_zero());
        return ok ? WIRE_OK : WIRE_NOT_REDUCED;
    } else if (kind == 1) {
        Jac<FqOps> p;
        int st = g1_decode(rec, p);
        store_fp(img, p.x); store_fp(img + 4, p.y); store_fp(img + 8, p.z);
        return st;
    }
    Jac<Fq2Ops> p;
    int st = g2_decode(rec, p);
    store_fp2(img, p.x); store_fp2(img + 8, p.y); store_fp2(img + 16, p.z);
    return st;
}
}
