// bn.hpp -- header-only C++ host layer over the C ABI (bn_b200.h), mirroring the `bn` crate's public API.
//
// The reference's host language is Rust; rustc/cargo are not available in the build image, so the host side
// above the C ABI is written in C++ (the Rust shim a maintainer would add is in INTEGRATION.md).  Names,
// argument meaning and error behaviour follow reference src/lib.rs:
//
//   bn::Fr                     src/lib.rs:15-54      (carried as its Montgomery image; arithmetic stays on the host crate)
//   bn::G1, bn::G2  operator*  src/lib.rs:79-163     `impl Mul<Fr>`
//   bn::Gt  pow(), operator*   src/lib.rs:165-179
//   bn::pairing(G1, G2) -> Gt  src/lib.rs:181-183
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

struct Fr {
    bn_fr v;
    bool operator==(const Fr& o) const { return std::memcmp(&v, &o.v, sizeof v) == 0; }
};
struct G1 {
    bn_g1 v;
};
struct G2 {
    bn_g2 v;
};
struct Gt {
    bn_gt v;
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
