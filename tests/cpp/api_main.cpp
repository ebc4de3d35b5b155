// C++ host-API check (include/bn.hpp over the C ABI).  Reads a vector file written by tests/test_cpp_api.py:
//   u64 n | n x G1 | n x G2 | n x Fr | n x Gt expected pairing | n x G1 expected (g1*fr) | n x Gt expected (pairing^fr)
// and replays the reference's usage pattern (pairing(p, q), p * s, gt.pow(s), gt * gt; cf. examples/joux.rs:19-21,
// src/groups/mod.rs:798-823).  Exit code 0 = every result bit-identical.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/bn.hpp"

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: api_main vectors.bin [--link-only]\n"); return 2; }
    if (argc > 2) { printf("link ok\n"); return 0; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    uint64_t n;
    if (fread(&n, 8, 1, f) != 1) return 2;
    auto g1 = rd<bn::G1>(f, n);
    auto g2 = rd<bn::G2>(f, n);
    auto fr = rd<bn::Fr>(f, n);
    auto want_gt = rd<bn::Gt>(f, n);
    auto want_g1 = rd<bn::G1>(f, n);
    auto want_pow = rd<bn::Gt>(f, n);
    fclose(f);
    try {
        bn::init(0);
        auto gt = bn::pairing_batch(g1, g2);
        for (size_t i = 0; i < n; i++)
            if (gt[i] != want_gt[i]) { fprintf(stderr, "pairing_batch mismatch at %zu\n", i); return 1; }
        // single-element API, as a bn-crate user would write it
        for (size_t i = 0; i < 3 && i < n; i++) {
            bn::Gt e = bn::pairing(g1[i], g2[i]);
            if (e != want_gt[i]) { fprintf(stderr, "pairing mismatch at %zu\n", i); return 1; }
            bn::G1 sp = g1[i] * fr[i];
            if (memcmp(&sp, &want_g1[i], sizeof sp)) { fprintf(stderr, "G1*Fr mismatch at %zu\n", i); return 1; }
            if (e.pow(fr[i]) != want_pow[i]) { fprintf(stderr, "Gt::pow mismatch at %zu\n", i); return 1; }
            if (bn::pairing(sp, g2[i]) != want_pow[i]) { fprintf(stderr, "bilinearity mismatch at %zu\n", i); return 1; }
        }
        auto fused = bn::pairing_pow_batch(g1, g2, fr);
        for (size_t i = 0; i < n; i++)
            if (fused[i] != want_pow[i]) { fprintf(stderr, "pairing_pow_batch mismatch at %zu\n", i); return 1; }
        if (gt[0].inverse() * gt[0] != bn::pairing(g1[1], g2[1]).pow(fr[1]).inverse() * want_pow[1]) {
            fprintf(stderr, "Gt::inverse mismatch\n");
            return 1;
        }
        auto pw = bn::pow_batch(gt, fr);
        auto mg = bn::mul_batch(g1, fr);
        for (size_t i = 0; i < n; i++) {
            if (pw[i] != want_pow[i]) { fprintf(stderr, "pow_batch mismatch at %zu\n", i); return 1; }
            if (memcmp(&mg[i], &want_g1[i], sizeof(bn::G1))) { fprintf(stderr, "mul_batch mismatch at %zu\n", i); return 1; }
        }
        // wire format round trip (src/groups/mod.rs:143-205): decode(encode(p)) is the same group element
        auto w1 = bn::encode_batch(g1);
        auto w2 = bn::encode_batch(g2);
        auto d1 = bn::decode_g1_batch(w1);
        auto d2 = bn::decode_g2_batch(w2);
        if (bn::encode_batch(d1) != w1 || bn::encode_batch(d2) != w2) { fprintf(stderr, "wire round trip mismatch\n"); return 1; }
        auto gt2 = bn::pairing_batch(d1, d2);
        for (size_t i = 0; i < n; i++)
            if (gt2[i] != want_gt[i]) { fprintf(stderr, "pairing of decoded points mismatch at %zu\n", i); return 1; }
        bool rejected = false;
        try {
            auto bad = w1;
            bad[0][bad[0].size() - 1] ^= 1;  // y off by one bit: "point is not on the curve"
            (void)bn::decode_g1_batch(bad);
        } catch (const bn::DecodeError& e) {
            rejected = e.index == 0 && (e.status == 3 || e.status == 2);
        }
        if (!rejected) { fprintf(stderr, "corrupted record was not rejected\n"); return 1; }
    } catch (const bn::Error& e) {
        fprintf(stderr, "bn::Error %d: %s\n", e.code, e.what());
        return 3;
    }
    printf("cpp api ok: %llu pairs\n", (unsigned long long)n);
    return 0;
}
