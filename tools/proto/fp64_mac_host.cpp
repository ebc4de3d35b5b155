// host harness of the FP64 multiply-accumulate prototype: rounds x (a[r] * b[r]) accumulated, result mod 2^544
#include "fp64_mac.cuh"
extern "C" void fp64_mac_rounds(const uint32_t* a, const uint32_t* b, int rounds, uint32_t* out17) {
    fp64proto::Cols c;
    fp64proto::cols_zero(c);
    for (int r = 0; r < rounds; r++) fp64proto::mac(c, fp64proto::to_d6(a + 8 * r), fp64proto::to_d6(b + 8 * r));
    fp64proto::cols_to_limbs(c, out17);
}
