// bn.hpp -- header-only C++ host layer over the C ABI (bn_b200.h), mirroring the `bn` crate's public API.
//
// The reference's host language is Rust; rustc/cargo are not available in the build image, so the host side
// above the C ABI is written in C++ (the Rust shim a maintainer would add is in INTEGRATION.md).  Names,
// argument meaning and error behaviour follow reference src/lib.rs:
//
//   bn::Fr   zero one pow inverse is_zero  + - * unary-   src/lib.rs:15-54   (carried as its Montgomery image)
//   bn::G1, bn::G2   zero one is_zero normalize  + - unary- * Fr  ==   src/lib.rs:56-163   (trait Group + operator impls)
//   bn::Gt   one pow inverse  *                                     src/lib.rs:165-179
//   bn::pairing(G1, G2) -> Gt                                       src/lib.rs:181-183
// Not mirrored: random() (RNG), from_str (decimal parsing), interpret -- host-only conveniences off the hot path.
//
// plus the batch forms the GPU exists for (pairing_batch, mul_batch, pow_batch).  Semantic "errors" do not
// exist on this path (infinity => Gt::one(), like src/groups/mod.rs:765-766); infrastructure failures (no GPU,
// CUDA error) throw bn::Error -- there is no CPU fallback.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "bn_b200.h"

namespace bn {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != 0) throw Error(rc, std::string("bn_b200: ") + bn_b200_last_error());
}
inline void init(int device = 0) { check(bn_b200_init(device)); }

inline void init_multi(int n_gpus = 0) { check(bn_b200_init_multi(n_gpus)); }  // all GPUs of the box behind one process

struct Fr {
    bn_fr v;
    static Fr zero() { return Fr{}; }                      // src/lib.rs:20
    static Fr one() {                                       // src/lib.rs:21
        Fr r;
        check(bn_b200_fr_one(&r.v));
        return r;
    }
    bool is_zero() const { return !(v.l[0] | v.l[1] | v.l[2] | v.l[3]); }  // src/lib.rs:27
    bool operator==(const Fr& o) const { return std::memcmp(&v, &o.v, sizeof v) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    Fr pow(const Fr& e) const {                             // src/lib.rs:24
        Fr r;
        check(bn_b200_fr_pow_batch(&v, &e.v, &r.v, 1));
        return r;
    }
    // Fr::inverse returns Option<Fr> (src/lib.rs:26): `ok` is false for zero
    Fr inverse(bool* ok = nullptr) const {
        if (ok) *ok = !is_zero();
        Fr r;
        check(bn_b200_fr_op_batch(4, &v, nullptr, &r.v, 1));
        return r;
    }
};
inline Fr fr_op(int op, const Fr& a, const Fr* b) {
    Fr r;
    check(bn_b200_fr_op_batch(op, &a.v, b ? &b->v : nullptr, &r.v, 1));
    return r;
}
inline Fr operator*(const Fr& a, const Fr& b) { return fr_op(0, a, &b); }  // src/lib.rs:50-54
inline Fr operator+(const Fr& a, const Fr& b) { return fr_op(1, a, &b); }  // src/lib.rs:32-36
inline Fr operator-(const Fr& a, const Fr& b) { return fr_op(2, a, &b); }  // src/lib.rs:38-42
inline Fr operator-(const Fr& a) { return fr_op(3, a, nullptr); }          // src/lib.rs:44-48

// trait Group (src/lib.rs:56-77) for G1 / G2
struct G1 {
    bn_g1 v;
    static G1 zero() { G1 r; check(bn_b200_g1_zero(&r.v)); return r; }
    static G1 one() { G1 r; check(bn_b200_g1_one(&r.v)); return r; }
    bool is_zero() const { return !(v.z[0] | v.z[1] | v.z[2] | v.z[3]); }
    void normalize() { bn_g1 t = v; check(bn_b200_g1_normalize_batch(&t, &v, 1)); }   // src/lib.rs:88-95
    bool operator==(const G1& o) const { uint8_t e = 0; check(bn_b200_g1_eq_batch(&v, &o.v, &e, 1)); return e != 0; }
    bool operator!=(const G1& o) const { return !(*this == o); }
};
struct G2 {
    bn_g2 v;
    static G2 zero() { G2 r; check(bn_b200_g2_zero(&r.v)); return r; }
    static G2 one() { G2 r; check(bn_b200_g2_one(&r.v)); return r; }
    bool is_zero() const {
        uint64_t o = 0;
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 4; j++) o |= v.z[i][j];
        return o == 0;
    }
    void normalize() { bn_g2 t = v; check(bn_b200_g2_normalize_batch(&t, &v, 1)); }   // src/lib.rs:131-138
    bool operator==(const G2& o) const { uint8_t e = 0; check(bn_b200_g2_eq_batch(&v, &o.v, &e, 1)); return e != 0; }
    bool operator!=(const G2& o) const { return !(*this == o); }
};
struct Gt {
    bn_gt v;
    static Gt one() { Gt r; check(bn_b200_gt_one(&r.v)); return r; }  // src/lib.rs:170
    bool operator==(const Gt& o) const { return std::memcmp(&v, &o.v, sizeof v) == 0; }
    bool operator!=(const Gt& o) const { return !(*this == o); }
    Gt inverse() const {  // Gt::inverse, src/lib.rs:172
        Gt r;
        check(bn_b200_gt_inv_batch(&v, &r.v, 1));
        return r;
    }
    Gt pow(const Fr& k) const {  // Gt::pow, src/lib.rs:171
        Gt r;
        check(bn_b200_gt_pow_batch(&v, &k.v, &r.v, 1));
        return r;
    }
};
static_assert(sizeof(Fr) == 32 && sizeof(G1) == 96 && sizeof(G2) == 192 && sizeof(Gt) == 384,
              "layouts must match the crate's #[repr(C)] types");

inline G1 g1_op(int op, const G1& a, const G1* b) {
    G1 r;
    check(bn_b200_g1_op_batch(op, &a.v, b ? &b->v : nullptr, &r.v, 1));
    return r;
}
inline G2 g2_op(int op, const G2& a, const G2* b) {
    G2 r;
    check(bn_b200_g2_op_batch(op, &a.v, b ? &b->v : nullptr, &r.v, 1));
    return r;
}
inline G1 operator+(const G1& a, const G1& b) { return g1_op(0, a, &b); }  // src/lib.rs:97-101
inline G1 operator-(const G1& a, const G1& b) { return g1_op(1, a, &b); }  // src/lib.rs:103-107
inline G1 operator-(const G1& a) { return g1_op(2, a, nullptr); }          // src/lib.rs:109-113
inline G2 operator+(const G2& a, const G2& b) { return g2_op(0, a, &b); }  // src/lib.rs:140-144
inline G2 operator-(const G2& a, const G2& b) { return g2_op(1, a, &b); }  // src/lib.rs:146-150
inline G2 operator-(const G2& a) { return g2_op(2, a, nullptr); }          // src/lib.rs:152-156
inline G1 operator*(const G1& p, const Fr& k) {  // src/lib.rs:116-120
    G1 r;
    check(bn_b200_g1_mul_batch(&p.v, &k.v, &r.v, 1));
    return r;
}
inline G2 operator*(const G2& p, const Fr& k) {  // src/lib.rs:159-163
    G2 r;
    check(bn_b200_g2_mul_batch(&p.v, &k.v, &r.v, 1));
    return r;
}
inline Gt operator*(const Gt& a, const Gt& b) {  // src/lib.rs:175-179
    Gt r;
    check(bn_b200_gt_mul_batch(&a.v, &b.v, &r.v, 1));
    return r;
}
inline Gt pairing(const G1& p, const G2& q) {  // src/lib.rs:181-183
    Gt r;
    check(bn_b200_pairing_batch(&p.v, &q.v, &r.v, 1));
    return r;
}

// ---- batch forms: one call, one H2D / kernels / D2H round trip ----
inline std::vector<Gt> pairing_batch(const std::vector<G1>& p, const std::vector<G2>& q) {
    if (p.size() != q.size()) throw Error(BN_B200_EINVAL, "pairing_batch: length mismatch");
    std::vector<Gt> out(p.size());
    check(bn_b200_pairing_batch(&p.data()->v, &q.data()->v, &out.data()->v, p.size()));
    return out;
}
// pairing(p[i], q[i]).pow(k[i]) fused on the device (the pattern of reference examples/joux.rs:19-21)
inline std::vector<Gt> pairing_pow_batch(const std::vector<G1>& p, const std::vector<G2>& q, const std::vector<Fr>& k) {
    if (p.size() != q.size() || p.size() != k.size()) throw Error(BN_B200_EINVAL, "pairing_pow_batch: length mismatch");
    std::vector<Gt> out(p.size());
    check(bn_b200_pairing_pow_batch(&p.data()->v, &q.data()->v, &k.data()->v, &out.data()->v, p.size()));
    return out;
}
inline std::vector<G1> mul_batch(const std::vector<G1>& p, const std::vector<Fr>& k) {
    if (p.size() != k.size()) throw Error(BN_B200_EINVAL, "mul_batch: length mismatch");
    std::vector<G1> out(p.size());
    check(bn_b200_g1_mul_batch(&p.data()->v, &k.data()->v, &out.data()->v, p.size()));
    return out;
}
inline std::vector<G2> mul_batch(const std::vector<G2>& p, const std::vector<Fr>& k) {
    if (p.size() != k.size()) throw Error(BN_B200_EINVAL, "mul_batch: length mismatch");
    std::vector<G2> out(p.size());
    check(bn_b200_g2_mul_batch(&p.data()->v, &k.data()->v, &out.data()->v, p.size()));
    return out;
}
// element-wise group law / Fr arithmetic on whole vectors (op codes of bn_b200.h)
inline std::vector<G1> add_batch(const std::vector<G1>& a, const std::vector<G1>& b) {
    if (a.size() != b.size()) throw Error(BN_B200_EINVAL, "add_batch: length mismatch");
    std::vector<G1> out(a.size());
    check(bn_b200_g1_add_batch(&a.data()->v, &b.data()->v, &out.data()->v, a.size()));
    return out;
}
inline std::vector<G2> add_batch(const std::vector<G2>& a, const std::vector<G2>& b) {
    if (a.size() != b.size()) throw Error(BN_B200_EINVAL, "add_batch: length mismatch");
    std::vector<G2> out(a.size());
    check(bn_b200_g2_add_batch(&a.data()->v, &b.data()->v, &out.data()->v, a.size()));
    return out;
}
inline std::vector<Fr> fr_batch(int op, const std::vector<Fr>& a, const std::vector<Fr>& b) {
    if (op <= 2 && a.size() != b.size()) throw Error(BN_B200_EINVAL, "fr_batch: length mismatch");
    std::vector<Fr> out(a.size());
    check(bn_b200_fr_op_batch(op, &a.data()->v, op <= 2 ? &b.data()->v : nullptr, &out.data()->v, a.size()));
    return out;
}
inline std::vector<Gt> pow_batch(const std::vector<Gt>& a, const std::vector<Fr>& k) {
    if (a.size() != k.size()) throw Error(BN_B200_EINVAL, "pow_batch: length mismatch");
    std::vector<Gt> out(a.size());
    check(bn_b200_gt_pow_batch(&a.data()->v, &k.data()->v, &out.data()->v, a.size()));
    return out;
}

// ---- wire format (RustcEncodable / RustcDecodable of G1, G2: src/groups/mod.rs:143-205) ----
// encode: one byte string per point, exactly the reference's (0x00 for infinity, else 0x04 | x | y).
// decode: throws bn::DecodeError with the reference's message for the first invalid element.
struct DecodeError : std::runtime_error {
    size_t index;
    int status;
    DecodeError(size_t i, int st)
        : std::runtime_error(st == 1   ? "invalid leading byte for uncompressed group element"
                             : st == 2 ? "integer is not less than modulus"
                             : st == 3 ? "point is not on the curve"
                                       : "point is not in the subgroup"),
          index(i), status(st) {}
};
namespace detail {
template <class P, class Raw, int REC>
std::vector<std::string> encode(const std::vector<P>& p, int (*fn)(const Raw*, uint8_t*, size_t)) {
    std::vector<uint8_t> rec(p.size() * REC);
    check(fn(&p.data()->v, rec.data(), p.size()));
    std::vector<std::string> out(p.size());
    for (size_t i = 0; i < p.size(); i++) {
        const char* r = reinterpret_cast<const char*>(rec.data() + i * REC);
        out[i].assign(r, r[0] == 0 ? 1 : REC);
    }
    return out;
}
template <class P, class Raw, int REC>
std::vector<P> decode(const std::vector<std::string>& w, int (*fn)(const uint8_t*, Raw*, uint8_t*, size_t)) {
    std::vector<uint8_t> rec(w.size() * REC, 0), st(w.size());
    for (size_t i = 0; i < w.size(); i++) {
        if (w[i].empty() || w[i].size() > (size_t)REC || (w[i][0] != 0 && w[i].size() != (size_t)REC)) throw DecodeError(i, 1);
        std::memcpy(rec.data() + i * REC, w[i].data(), w[i].size());
    }
    std::vector<P> out(w.size());
    check(fn(rec.data(), &out.data()->v, st.data(), w.size()));
    for (size_t i = 0; i < w.size(); i++)
        if (st[i]) throw DecodeError(i, st[i]);
    return out;
}
}  // namespace detail
inline std::vector<std::string> encode_batch(const std::vector<G1>& p) {
    return detail::encode<G1, bn_g1, BN_B200_G1_WIRE_BYTES>(p, bn_b200_g1_encode_batch);
}
inline std::vector<std::string> encode_batch(const std::vector<G2>& p) {
    return detail::encode<G2, bn_g2, BN_B200_G2_WIRE_BYTES>(p, bn_b200_g2_encode_batch);
}
inline std::vector<G1> decode_g1_batch(const std::vector<std::string>& w) {
    return detail::decode<G1, bn_g1, BN_B200_G1_WIRE_BYTES>(w, bn_b200_g1_decode_batch);
}
inline std::vector<G2> decode_g2_batch(const std::vector<std::string>& w) {
    return detail::decode<G2, bn_g2, BN_B200_G2_WIRE_BYTES>(w, bn_b200_g2_decode_batch);
}

}  // namespace bn
