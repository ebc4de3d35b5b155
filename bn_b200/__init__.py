"""bn_b200 -- B200-native batched BN254 engine behind the `bn` crate's public API.

Host-side mirror (Python, for tests and bench; the C++/Rust bindings are in include/ and INTEGRATION.md) of
reference src/lib.rs: Fr, G1, G2, Gt, pairing().  Everything numeric is computed by libbn_b200.so (hand-written
sm_100a kernels); importing this package without the built library, or calling it without a GPU, raises.
"""
from ._lib import BnB200Error, init, init_multi, load, pairing_kernel_names, shutdown  # noqa: F401
from .api import (Fr, G1, G2, Gt, fq_sqr_chain, fr_pow_batch, g1_eq_batch, g1_op_batch, g2_eq_batch, g2_op_batch, gt_exp_by_neg_z_batch, decode_batch, encode_batch, fq_mul_chain, fr_op_batch, g1_check_batch, g2_check_batch, g1_normalize_batch, g2_normalize_batch, g1_mul_batch, g2_mul_batch, gt_inv_batch, gt_mul_batch, gt_pow_batch,  # noqa: F401
                  pairing, pairing_batch, pairing_pow_batch, to_wire, from_wire)
