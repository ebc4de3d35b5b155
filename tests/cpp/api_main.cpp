// C++ host-API check (include/bn.hpp over the C ABI).
//
//   api_main vectors.bin            replay the reference's usage pattern on oracle-generated vectors (tests/test_cpp_api.py):
//       u64 n | n x G1 | n x G2 | n x Fr | n x Gt pairing | n x G1 (g1*fr) | n x Gt (pairing^fr)
//             | n x G1 (g1[i] + g1[(i+1)%n]) | n x G2 (g2[i] - g2[(i+1)%n]) | n x G1 (g1.double()) | n x Fr (fr[i]^fr[(i+1)%n])
//       (pairing(p, q), p * s, gt.pow(s), gt * gt, p + q, -p, Fr arithmetic; cf. examples/joux.rs:19-21,
//       src/groups/mod.rs:798-823, src/groups/tests.rs).  Exit code 0 = every result bit-identical.
//   api_main --multi G PAIRS out.bin    ONE process drives G GPUs (bn::init_multi): PAIRS pairings through the host-pointer
//       call, sharded inside the library; compares with the single-GPU results, times both, and writes a sample
//       (u64 m | m x G1 | m x G2 | m x Gt) for the caller to check against the oracle.
//   api_main /dev/null --link-only  link check (CPU test)
#include <chrono>
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
template <class T>
static bool same(const T& a, const T& b) { return memcmp(&a, &b, sizeof(T)) == 0; }

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int run_multi(int gpus, size_t pairs, const char* out_path) {
    using namespace bn;
    init_multi(gpus);
    const int bound = bn_b200_device_count();
    // page-locked buffers from the library (no CUDA runtime linked here)
    bn_g1* p; bn_g2* q; bn_gt *o_multi, *o_one; bn_fr *ka, *kb; bn_g1* base1; bn_g2* base2;
    check(bn_b200_alloc_pinned((void**)&p, pairs * sizeof(bn_g1)));
    check(bn_b200_alloc_pinned((void**)&q, pairs * sizeof(bn_g2)));
    check(bn_b200_alloc_pinned((void**)&o_multi, pairs * sizeof(bn_gt)));
    check(bn_b200_alloc_pinned((void**)&o_one, pairs * sizeof(bn_gt)));
    check(bn_b200_alloc_pinned((void**)&ka, pairs * sizeof(bn_fr)));
    check(bn_b200_alloc_pinned((void**)&kb, pairs * sizeof(bn_fr)));
    check(bn_b200_alloc_pinned((void**)&base1, pairs * sizeof(bn_g1)));
    check(bn_b200_alloc_pinned((void**)&base2, pairs * sizeof(bn_g2)));
    // inputs: P_i = G1::one() * a_i, Q_i = G2::one() * b_i (the multi-GPU scalar-mul entry points build them);
    // a_i, b_i = splitmix64 words with the top limb masked below r's: valid canonical Fr images
    uint64_t st = 0xB2000005ULL;
    auto next = [&]() { st += 0x9E3779B97F4A7C15ULL; uint64_t z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); };
    G1 g1 = G1::one();
    G2 g2 = G2::one();
    for (size_t i = 0; i < pairs; i++) {
        for (int l = 0; l < 4; l++) { ka[i].l[l] = next(); kb[i].l[l] = next(); }
        ka[i].l[3] &= 0x0FFFFFFFFFFFFFFFULL;
        kb[i].l[3] &= 0x0FFFFFFFFFFFFFFFULL;
        base1[i] = g1.v;
        base2[i] = g2.v;
    }
    check(bn_b200_g1_mul_batch(base1, ka, p, pairs));
    check(bn_b200_g2_mul_batch(base2, kb, q, pairs));
    // multi-GPU: warm up (three calls: both staging sets of every device get allocated, modules are loaded), then time
    for (int r = 0; r < 3; r++) check(bn_b200_pairing_batch(p, q, o_multi, pairs));
    const int reps = 10;
    double t0 = now();
    for (int r = 0; r < reps; r++) check(bn_b200_pairing_batch(p, q, o_multi, pairs));
    const double multi_rate = reps * (double)pairs / (now() - t0);
    // one GPU: a single device's share of the batch per call (the per-GPU load of the run above), then the whole batch for comparison
    init(0);
    const size_t share = (pairs + bound - 1) / bound;
    for (int r = 0; r < 3; r++) check(bn_b200_pairing_batch(p, q, o_one, share));
    t0 = now();
    for (int r = 0; r < reps; r++) check(bn_b200_pairing_batch(p, q, o_one, share));
    const double one_rate = reps * (double)share / (now() - t0);
    check(bn_b200_pairing_batch(p, q, o_one, pairs));
    size_t bad = 0;
    for (size_t i = 0; i < pairs; i++) bad += !same(o_multi[i], o_one[i]);
    printf("multi: %d GPUs bound, %zu pairs: %.0f pairings/s; one GPU on its %zu-pair share: %.0f pairings/s; ratio %.3f (ideal %d); mismatches vs one GPU: %zu\n",
           bound, pairs, multi_rate, share, one_rate, multi_rate / one_rate, bound, bad);
    // sample for the oracle: evenly spaced + both sides of every shard boundary
    std::vector<size_t> idx;
    for (size_t i = 0; i < pairs; i += pairs / 96 ? pairs / 96 : 1) idx.push_back(i);
    for (int d = 1; d < bound; d++) {
        size_t b = ((pairs + bound - 1) / bound + 19) / 20 * 20 * d;
        if (b < pairs) { idx.push_back(b - 1); idx.push_back(b); }
    }
    idx.push_back(pairs - 1);
    FILE* f = fopen(out_path, "wb");
    if (!f) { perror("open"); return 2; }
    uint64_t m = idx.size();
    fwrite(&m, 8, 1, f);
    for (size_t i : idx) fwrite(&p[i], sizeof(bn_g1), 1, f);
    for (size_t i : idx) fwrite(&q[i], sizeof(bn_g2), 1, f);
    for (size_t i : idx) fwrite(&o_multi[i], sizeof(bn_gt), 1, f);
    fclose(f);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: api_main vectors.bin [--link-only] | --multi G PAIRS out.bin\n"); return 2; }
    if (argc > 2 && std::string(argv[2]) == "--link-only") { printf("link ok\n"); return 0; }
    try {
        if (std::string(argv[1]) == "--multi" && argc >= 5) return run_multi(atoi(argv[2]), (size_t)atoll(argv[3]), argv[4]);
    } catch (const bn::Error& e) {
        fprintf(stderr, "bn::Error %d: %s\n", e.code, e.what());
        return 3;
    }
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
    auto want_add = rd<bn::G1>(f, n);
    auto want_sub2 = rd<bn::G2>(f, n);
    auto want_dbl = rd<bn::G1>(f, n);
    auto want_frpow = rd<bn::Fr>(f, n);
    fclose(f);
#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)
    try {
        bn::init(0);
        auto gt = bn::pairing_batch(g1, g2);
        for (size_t i = 0; i < n; i++)
            if (gt[i] != want_gt[i]) FAIL("pairing_batch mismatch at %zu", i);
        // single-element API, as a bn-crate user would write it
        for (size_t i = 0; i < 3 && i < n; i++) {
            bn::Gt e = bn::pairing(g1[i], g2[i]);
            if (e != want_gt[i]) FAIL("pairing mismatch at %zu", i);
            bn::G1 sp = g1[i] * fr[i];
            if (!same(sp, want_g1[i])) FAIL("G1*Fr mismatch at %zu", i);
            if (e.pow(fr[i]) != want_pow[i]) FAIL("Gt::pow mismatch at %zu", i);
            if (bn::pairing(sp, g2[i]) != want_pow[i]) FAIL("bilinearity mismatch at %zu", i);
        }
        auto fused = bn::pairing_pow_batch(g1, g2, fr);
        for (size_t i = 0; i < n; i++)
            if (fused[i] != want_pow[i]) FAIL("pairing_pow_batch mismatch at %zu", i);
        if (gt[0].inverse() * gt[0] != bn::pairing(g1[1], g2[1]).pow(fr[1]).inverse() * want_pow[1]) FAIL("Gt::inverse mismatch");
        if (gt[0].inverse() * gt[0] != bn::Gt::one()) FAIL("Gt::one mismatch");
        auto pw = bn::pow_batch(gt, fr);
        auto mg = bn::mul_batch(g1, fr);
        for (size_t i = 0; i < n; i++) {
            if (pw[i] != want_pow[i]) FAIL("pow_batch mismatch at %zu", i);
            if (!same(mg[i], want_g1[i])) FAIL("mul_batch mismatch at %zu", i);
        }
        // group law and Fr arithmetic through the operators (src/lib.rs:19-54, 97-114, 140-157): limb-exact vs the oracle
        for (size_t i = 0; i < n; i++) {
            const size_t j = (i + 1) % n;
            if (!same(g1[i] + g1[j], want_add[i])) FAIL("G1 + mismatch at %zu", i);
            if (!same(g2[i] - g2[j], want_sub2[i])) FAIL("G2 - mismatch at %zu", i);
            if (!same(bn::g1_op(3, g1[i], nullptr), want_dbl[i])) FAIL("G1 double mismatch at %zu", i);
            if (fr[i].pow(fr[j]) != want_frpow[i]) FAIL("Fr::pow mismatch at %zu", i);
        }
        // group-law identities as the reference's group_trials check them (src/groups/tests.rs:5-102), projective ==
        {
            bn::G1 a = g1[0], b = g1[1], c = g1[2];
            if (!((a + b) + c == a + (b + c))) FAIL("G1 associativity");
            if (!(a + b == b + a)) FAIL("G1 commutativity");
            if (!((a - a).is_zero()) || !((a + (-a)).is_zero())) FAIL("G1 a - a != 0");
            if (!(a + bn::G1::zero() == a) || !bn::G1::zero().is_zero() || bn::G1::one().is_zero()) FAIL("G1 zero / one");
            if (!(a + a == bn::g1_op(3, a, nullptr))) FAIL("G1 a + a != double");
            bn::G1 an = a;
            an.normalize();
            if (!(an == a) || (!a.is_zero() && !same(an.v.z, bn::G1::one().v.z))) FAIL("G1 normalize");
            if (!(bn::G1::one() * fr[0] + bn::G1::one() * fr[1] == bn::G1::one() * (fr[0] + fr[1]))) FAIL("G1 distributivity");
            bn::G2 x = g2[0], y = g2[1];
            if (!(x + y == y + x) || !((x - x).is_zero()) || !(x + bn::G2::zero() == x)) FAIL("G2 group law");
            if (!(bn::G2::one() * fr[0] * fr[1] == bn::G2::one() * (fr[0] * fr[1]))) FAIL("G2 scalar associativity");
            bn::Fr s = fr[0], t = fr[1];
            bool ok = false;
            if ((s * t) * s.inverse(&ok) != t || !ok) FAIL("Fr (s*t)/s != t");
            if (s + (-s) != bn::Fr::zero() || s - s != bn::Fr::zero() || s * bn::Fr::one() != s) FAIL("Fr identities");
            if (s.pow(bn::Fr::one() + bn::Fr::one()) != s * s) FAIL("Fr pow");
            (void)bn::Fr::zero().inverse(&ok);
            if (ok) FAIL("Fr zero inverse must report failure");
        }
        // wire format round trip (src/groups/mod.rs:143-205): decode(encode(p)) is the same group element
        auto w1 = bn::encode_batch(g1);
        auto w2 = bn::encode_batch(g2);
        auto d1 = bn::decode_g1_batch(w1);
        auto d2 = bn::decode_g2_batch(w2);
        if (bn::encode_batch(d1) != w1 || bn::encode_batch(d2) != w2) FAIL("wire round trip mismatch");
        auto gt2 = bn::pairing_batch(d1, d2);
        for (size_t i = 0; i < n; i++)
            if (gt2[i] != want_gt[i]) FAIL("pairing of decoded points mismatch at %zu", i);
        bool rejected = false;
        try {
            auto bad = w1;
            bad[0][bad[0].size() - 1] ^= 1;  // y off by one bit: "point is not on the curve"
            (void)bn::decode_g1_batch(bad);
        } catch (const bn::DecodeError& e) {
            rejected = e.index == 0 && (e.status == 3 || e.status == 2);
        }
        if (!rejected) FAIL("corrupted record was not rejected");
    } catch (const bn::Error& e) {
        fprintf(stderr, "bn::Error %d: %s\n", e.code, e.what());
        return 3;
    }
    printf("cpp api ok: %llu pairs\n", (unsigned long long)n);
    return 0;
}
