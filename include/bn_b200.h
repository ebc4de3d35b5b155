/*
 * bn_b200.h -- C ABI of the B200-native batched BN254 engine (libbn_b200.so).
 *
 * This is the drop-in boundary for the `bn` crate (zcash-hackworks/bn): every struct below is
 * byte-identical to the crate's #[repr(C)] public type, so `&[G1]`, `&[G2]`, `&mut [Gt]` cross the FFI as
 * raw pointers with no marshalling.  All field elements are Montgomery form (x * 2^256 mod p), canonical
 * in [0, p), four little-endian u64 limbs -- exactly what the crate keeps in memory
 * (reference src/arith.rs:9-11, src/fields/fp.rs:11-13).
 *
 * The reference has no FFI of its own; each entry point names the public Rust item it replaces.
 * INTEGRATION.md shows the `extern "C"` block and the thin Rust wrappers a maintainer would add.
 *
 * Conventions
 *   - return 0 on success, a negative BN_B200_E* code on failure; never unwinds, never aborts;
 *     bn_b200_last_error() gives a message for the last failure on the calling thread.
 *   - caller owns every buffer; the library owns its stream, scratch memory and events.
 *   - host-pointer entry points copy H2D, run the kernels and copy D2H before returning (bn_b200_pairing_batch writes
 *     a page-locked, device-mapped `out` buffer directly from the last kernel instead of a D2H pass).  With several
 *     GPUs bound (bn_b200_init_multi) a host-pointer batch is split into contiguous ranges, one per GPU, inside the call.
 *   - *_dev entry points take DEVICE pointers (same layouts) and enqueue on `stream` without synchronising.  `stream` is a
 *     cudaStream_t cast to void*.  NULL is NOT the caller's default stream: it selects the library's own non-blocking
 *     stream of that device (pass cudaStreamLegacy / cudaStreamPerThread explicitly to use a default stream).  The call runs
 *     on the bound device that owns the pointers.  Calls on different streams are safe: the library orders its own scratch
 *     memory between them with events; the caller orders its own buffers.
 *   - thread-safe (the crate's types are Send + Sync, reference src/lib.rs:56-66): one lock per device, held while work is
 *     enqueued, released before the host waits; the CUDA device of the calling thread is saved and restored by every call.
 *   - there is NO CPU fallback: without a CUDA device every call fails with BN_B200_ENODEV.
 */
#ifndef BN_B200_H
#define BN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } bn_fr;                          /* bn::Fr  src/lib.rs:15-17     32 B */
typedef struct { uint64_t x[4], y[4], z[4]; } bn_g1;              /* bn::G1  src/lib.rs:79-81     96 B, Jacobian */
typedef struct { uint64_t x[2][4], y[2][4], z[2][4]; } bn_g2;     /* bn::G2  src/lib.rs:122-124  192 B, Fq2 = c0 then c1 */
typedef struct { uint64_t c[2][3][2][4]; } bn_gt;                 /* bn::Gt  src/lib.rs:165-167  384 B, Fq12.c{0,1}.c{0,1,2}.c{0,1} */

#define BN_B200_OK 0
#define BN_B200_ENODEV (-1)   /* no CUDA device / driver */
#define BN_B200_ECUDA (-2)    /* a CUDA runtime call failed */
#define BN_B200_EINVAL (-3)   /* null pointer or bad argument */
#define BN_B200_ENOMEM (-4)   /* device allocation failed */

/* Bind ONE CUDA device to this process (one process per GPU) and create its streams.  Idempotent; binding a different
 * device tears the previous binding down first (streams, scratch and staging memory are released). */
int bn_b200_init(int device);
/* Bind devices 0 .. n_gpus-1 (n_gpus <= 0: every visible device, at most 8) to THIS process: the library then owns one
 * context (streams, scratch, staging) per GPU and host-pointer batches are sharded over all of them from one call
 * ("bn_b200_init(int n_gpus) creates streams").  No torch / NCCL involved: shards are
 * independent, results land in the caller's buffer (zero-copy stores into page-locked memory, or one D2H per device). */
int bn_b200_init_multi(int n_gpus);
/* Number of devices currently bound (0 before init). */
int bn_b200_device_count(void);
int bn_b200_shutdown(void);
/* A device is only given a shard of at least `elements` (default 2048; 0 restores it): small batches stay on one GPU. */
int bn_b200_set_min_shard(size_t elements);
/* Page-locked host memory mapped into every bound device, for callers that do not link the CUDA runtime themselves:
 * inputs in it are copied without staging, and an `out` buffer in it is written directly by the last pairing kernel. */
int bn_b200_alloc_pinned(void** p, size_t bytes);
int bn_b200_free_pinned(void* p);
/* OR of the bound devices' error words (bit 0: a line-ring / TMA wait timed out inside a pairing kernel: results of that
 * call are invalid).  Host-pointer calls check it themselves and return BN_B200_ECUDA; *_dev callers poll it here.
 * Synchronises the devices.  clear != 0 resets the words.  Negative: BN_B200_E*. */
int bn_b200_device_error(int clear);
const char* bn_b200_last_error(void);
/* Number of SMs of the active device (0 before init). */
int bn_b200_sm_count(void);
/* Line evaluations per pairing in the library's Miller schedule (88: NAF walk of 6u+2; the reference's binary walk has 102). */
int bn_b200_num_lines(void);

/* Constants of the crate's API as byte images (pure data; no device needed).
 * replaces Fr::one src/lib.rs:21, Group::one / zero src/lib.rs:84-85, 127-128 (src/groups/mod.rs:208-218, 356-390), Gt::one src/lib.rs:170.
 * (Fr::zero is all-zero bytes.) */
int bn_b200_fr_one(bn_fr* out);
int bn_b200_g1_one(bn_g1* out);
int bn_b200_g2_one(bn_g2* out);
int bn_b200_g1_zero(bn_g1* out);
int bn_b200_g2_zero(bn_g2* out);
int bn_b200_gt_one(bn_gt* out);

/* pairing(p, q) for n independent pairs.              replaces bn::pairing, src/lib.rs:181-183
 * (groups::pairing src/groups/mod.rs:764-771: to_affine + precompute + miller_loop + final_exponentiation;
 *  either input at infinity => Gt::one()). */
int bn_b200_pairing_batch(const bn_g1* p, const bn_g2* q, bn_gt* out, size_t n);
int bn_b200_pairing_batch_dev(const bn_g1* d_p, const bn_g2* d_q, bn_gt* d_out, size_t n, void* stream);

/* Multi-GPU form: this rank's n results are stored by the final-exponentiation kernel's epilogue
 * directly into EVERY rank's gather buffer, peer_out[r][rank * n + i] for r = 0..world-1 (peer-mapped device memory
 * reachable over NVLink, e.g. a CUDA-IPC / symmetric-memory allocation; peer_out is a HOST array of `world` device
 * pointers, each to world * n elements).  Replaces pairing_batch_dev + an all-gather of Gt; the caller synchronises the
 * ranks (barrier) before reading.  world <= 8. */
int bn_b200_pairing_batch_gather_dev(const bn_g1* d_p, const bn_g2* d_q, bn_gt* const* peer_out, int world, int rank, size_t n,
                                     void* stream);

/* out[i] = pairing(p[i], q[i]).pow(k[i]) in one pass (the pattern of reference examples/joux.rs:19-21
 * and test_binlinearity, src/groups/mod.rs:811).  Same bytes as bn_b200_pairing_batch followed by bn_b200_gt_pow_batch. */
int bn_b200_pairing_pow_batch(const bn_g1* p, const bn_g2* q, const bn_fr* k, bn_gt* out, size_t n);
int bn_b200_pairing_pow_batch_dev(const bn_g1* d_p, const bn_g2* d_q, const bn_fr* d_k, bn_gt* d_out, size_t n, void* stream);

/* out[i] = p[i] * k[i].                               replaces `impl Mul<Fr> for G1/G2`, src/lib.rs:116-120, 159-163
 * Output is the same un-normalised Jacobian triple the crate produces (src/groups/mod.rs:250-270). */
int bn_b200_g1_mul_batch(const bn_g1* p, const bn_fr* k, bn_g1* out, size_t n);
int bn_b200_g1_mul_batch_dev(const bn_g1* d_p, const bn_fr* d_k, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g2_mul_batch(const bn_g2* p, const bn_fr* k, bn_g2* out, size_t n);
int bn_b200_g2_mul_batch_dev(const bn_g2* d_p, const bn_fr* d_k, bn_g2* d_out, size_t n, void* stream);

/* Group law at the boundary.       replaces `impl Add / Sub / Neg for G1, G2`, src/lib.rs:97-114, 140-157
 * (-> src/groups/mod.rs:272-347: add returns the other operand when one is zero, doubles when the operands are equal, and
 * P + (-P) falls through the general formula to a z = 0 triple; neg leaves zero unchanged; sub = a + (-b)) and G::double
 * (src/groups/mod.rs:228-247).  Outputs are the crate's un-normalised Jacobian triples, limb for limb.
 * bn_b200_g{1,2}_op_batch: op 0 a + b, 1 a - b, 2 -a, 3 a.double() (b may be NULL for op >= 2). */
int bn_b200_g1_op_batch(int op, const bn_g1* a, const bn_g1* b, bn_g1* out, size_t n);
int bn_b200_g1_op_batch_dev(int op, const bn_g1* d_a, const bn_g1* d_b, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g2_op_batch(int op, const bn_g2* a, const bn_g2* b, bn_g2* out, size_t n);
int bn_b200_g2_op_batch_dev(int op, const bn_g2* d_a, const bn_g2* d_b, bn_g2* d_out, size_t n, void* stream);
int bn_b200_g1_add_batch(const bn_g1* a, const bn_g1* b, bn_g1* out, size_t n);
int bn_b200_g1_add_batch_dev(const bn_g1* d_a, const bn_g1* d_b, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g1_sub_batch(const bn_g1* a, const bn_g1* b, bn_g1* out, size_t n);
int bn_b200_g1_sub_batch_dev(const bn_g1* d_a, const bn_g1* d_b, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g1_neg_batch(const bn_g1* a, bn_g1* out, size_t n);
int bn_b200_g1_neg_batch_dev(const bn_g1* d_a, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g1_double_batch(const bn_g1* a, bn_g1* out, size_t n);
int bn_b200_g1_double_batch_dev(const bn_g1* d_a, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g2_add_batch(const bn_g2* a, const bn_g2* b, bn_g2* out, size_t n);
int bn_b200_g2_add_batch_dev(const bn_g2* d_a, const bn_g2* d_b, bn_g2* d_out, size_t n, void* stream);
int bn_b200_g2_sub_batch(const bn_g2* a, const bn_g2* b, bn_g2* out, size_t n);
int bn_b200_g2_sub_batch_dev(const bn_g2* d_a, const bn_g2* d_b, bn_g2* d_out, size_t n, void* stream);
int bn_b200_g2_neg_batch(const bn_g2* a, bn_g2* out, size_t n);
int bn_b200_g2_neg_batch_dev(const bn_g2* d_a, bn_g2* d_out, size_t n, void* stream);
int bn_b200_g2_double_batch(const bn_g2* a, bn_g2* out, size_t n);
int bn_b200_g2_double_batch_dev(const bn_g2* d_a, bn_g2* d_out, size_t n, void* stream);
/* eq[i] = (a[i] == b[i]) as group elements.            replaces `PartialEq for G` (projective), src/groups/mod.rs:83-109 */
int bn_b200_g1_eq_batch(const bn_g1* a, const bn_g1* b, uint8_t* eq, size_t n);
int bn_b200_g1_eq_batch_dev(const bn_g1* d_a, const bn_g1* d_b, uint8_t* d_eq, size_t n, void* stream);
int bn_b200_g2_eq_batch(const bn_g2* a, const bn_g2* b, uint8_t* eq, size_t n);
int bn_b200_g2_eq_batch_dev(const bn_g2* d_a, const bn_g2* d_b, uint8_t* d_eq, size_t n, void* stream);

/* out[i] = a[i].pow(k[i]).                            replaces Gt::pow, src/lib.rs:171 (FieldElement::pow, src/fields/mod.rs:35-46) */
int bn_b200_gt_pow_batch(const bn_gt* a, const bn_fr* k, bn_gt* out, size_t n);
int bn_b200_gt_pow_batch_dev(const bn_gt* d_a, const bn_fr* d_k, bn_gt* d_out, size_t n, void* stream);
/* out[i] = a[i] * b[i].                               replaces `impl Mul<Gt> for Gt`, src/lib.rs:175-179 */
int bn_b200_gt_mul_batch(const bn_gt* a, const bn_gt* b, bn_gt* out, size_t n);
int bn_b200_gt_mul_batch_dev(const bn_gt* d_a, const bn_gt* d_b, bn_gt* d_out, size_t n, void* stream);

/* out[i] = a[i].inverse() (a[i] != 0).                 replaces Gt::inverse, src/lib.rs:172 */
int bn_b200_gt_inv_batch(const bn_gt* a, bn_gt* out, size_t n);
int bn_b200_gt_inv_batch_dev(const bn_gt* d_a, bn_gt* d_out, size_t n, void* stream);

/* out[i] = a[i].exp_by_neg_z() evaluated as the reference does (binary cyclotomic_pow(u), literal Granger-Scott squaring,
 * conjugate): defined for ANY Fq12 input.      replaces Fq12::exp_by_neg_z, src/fields/fq12.rs:97-101, 178-246
 * (the building block of final_exponentiation; exposed so the reference's test_cyclotomic_exp vector, src/fields/mod.rs:171-201,
 * runs on the device). */
int bn_b200_gt_exp_by_neg_z_batch(const bn_gt* a, bn_gt* out, size_t n);
int bn_b200_gt_exp_by_neg_z_batch_dev(const bn_gt* d_a, bn_gt* d_out, size_t n, void* stream);

/* out[i] = a[i].pow(e[i]).                             replaces Fr::pow, src/lib.rs:24 (FieldElement::pow, src/fields/mod.rs:35-46;
 * the exponent is U256::from(e[i])). */
int bn_b200_fr_pow_batch(const bn_fr* a, const bn_fr* e, bn_fr* out, size_t n);
int bn_b200_fr_pow_batch_dev(const bn_fr* d_a, const bn_fr* d_e, bn_fr* d_out, size_t n, void* stream);

/* Batched Fr arithmetic.            replaces bn::Fr Mul/Add/Sub/Neg/inverse, src/lib.rs:25, 32-54
 * op: 0 a*b, 1 a+b, 2 a-b, 3 -a, 4 a.inverse() (inverse of zero yields zero; the crate returns None).  b may be NULL for op >= 3. */
int bn_b200_fr_op_batch(int op, const bn_fr* a, const bn_fr* b, bn_fr* out, size_t n);
int bn_b200_fr_op_batch_dev(int op, const bn_fr* d_a, const bn_fr* d_b, bn_fr* d_out, size_t n, void* stream);

/* Group::normalize for n points (row f-3: the device half of wire encoding: affine x, y with z = one; infinity unchanged).
 * replaces src/lib.rs:88-95, 131-138 (G::to_affine, src/groups/mod.rs:113-130). */
int bn_b200_g1_normalize_batch(const bn_g1* p, bn_g1* out, size_t n);
int bn_b200_g1_normalize_batch_dev(const bn_g1* d_p, bn_g1* d_out, size_t n, void* stream);
int bn_b200_g2_normalize_batch(const bn_g2* p, bn_g2* out, size_t n);
int bn_b200_g2_normalize_batch_dev(const bn_g2* d_p, bn_g2* d_out, size_t n, void* stream);

/* Decode-side validity of n points given as (x, y, z = one) (row f-3): ok[i] = 1 iff on the curve and, for G2, in the
 * order-r subgroup (p * (-1) + p == zero); z = 0 is accepted.  replaces the checks of AffineG::decode,
 * src/groups/mod.rs:178-205 (the byte parsing and the `< modulus` range checks stay on the host). */
int bn_b200_g1_check_batch(const bn_g1* p, uint8_t* ok, size_t n);
int bn_b200_g1_check_batch_dev(const bn_g1* d_p, uint8_t* d_ok, size_t n, void* stream);
int bn_b200_g2_check_batch(const bn_g2* p, uint8_t* ok, size_t n);
int bn_b200_g2_check_batch_dev(const bn_g2* d_p, uint8_t* d_ok, size_t n, void* stream);

/* Wire format (row f-3), fixed-stride records.  replaces the RustcEncodable / RustcDecodable impls of Fr, G1, G2
 * (src/lib.rs:15, 79, 122 -> src/fields/fp.rs:24-36, src/fields/fq2.rs:31-53, src/groups/mod.rs:143-205; bincode adds no framing).
 *   Fr : 32 bytes, big-endian canonical integer.
 *   G1 : BN_B200_G1_WIRE_BYTES = 65 per record: 0x04 | x (32, BE) | y (32, BE).
 *   G2 : BN_B200_G2_WIRE_BYTES = 129 per record: 0x04 | x (64, BE integer c1*q + c0) | y (64).
 * The point at infinity is the single byte 0x00 on the wire; inside a fixed-stride record it is 0x00 followed by zero
 * padding (encode) and only byte 0 is read (decode) -- a caller producing the exact reference stream emits record[0..1].
 * encode normalises (to_affine) first.  decode writes (x, y, one) in Montgomery form, or zero() = (0, 1, 0) when the
 * record is infinity or invalid, and a status per element:
 *   0 ok | 1 invalid leading byte | 2 integer is not less than modulus | 3 not on the curve | 4 not in the subgroup (G2)
 * -- the reference's Err cases, checked in the reference's order (src/groups/mod.rs:178-205). */
#define BN_B200_G1_WIRE_BYTES 65
#define BN_B200_G2_WIRE_BYTES 129
#define BN_B200_FR_WIRE_BYTES 32
int bn_b200_g1_encode_batch(const bn_g1* p, uint8_t* out, size_t n);
int bn_b200_g1_encode_batch_dev(const bn_g1* d_p, uint8_t* d_out, size_t n, void* stream);
int bn_b200_g2_encode_batch(const bn_g2* p, uint8_t* out, size_t n);
int bn_b200_g2_encode_batch_dev(const bn_g2* d_p, uint8_t* d_out, size_t n, void* stream);
int bn_b200_fr_encode_batch(const bn_fr* a, uint8_t* out, size_t n);
int bn_b200_fr_encode_batch_dev(const bn_fr* d_a, uint8_t* d_out, size_t n, void* stream);
int bn_b200_g1_decode_batch(const uint8_t* in, bn_g1* out, uint8_t* status, size_t n);
int bn_b200_g1_decode_batch_dev(const uint8_t* d_in, bn_g1* d_out, uint8_t* d_status, size_t n, void* stream);
int bn_b200_g2_decode_batch(const uint8_t* in, bn_g2* out, uint8_t* status, size_t n);
int bn_b200_g2_decode_batch_dev(const uint8_t* d_in, bn_g2* d_out, uint8_t* d_status, size_t n, void* stream);
int bn_b200_fr_decode_batch(const uint8_t* in, bn_fr* out, uint8_t* status, size_t n);
int bn_b200_fr_decode_batch_dev(const uint8_t* d_in, bn_fr* d_out, uint8_t* d_status, size_t n, void* stream);

/* x <- x * b (Montgomery, mod q) repeated `iters` times per element: the BASELINE config-2 microbenchmark of
 * the innermost operation (Fq Mul, src/fields/fp.rs:137-146 -> U256::mul src/arith.rs:257-263). a, b, out: n x 4 u64. */
int bn_b200_fq_mul_chain(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, uint32_t iters);
int bn_b200_fq_mul_chain_dev(const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, size_t n, uint32_t iters, void* stream);

/* x <- x^2 (Montgomery, mod q) repeated `iters` times: the dedicated squaring (108 IMAD.WIDE against 136). */
int bn_b200_fq_sqr_chain(const uint64_t* a, uint64_t* out, size_t n, uint32_t iters);
int bn_b200_fq_sqr_chain_dev(const uint64_t* d_a, uint64_t* d_out, size_t n, uint32_t iters, void* stream);

/* Calibration kernel for the roofline denominator: `blocks` x 256 threads each issue iters*32 independent
 * IMAD.WIDE.U32 (the instruction every Fq product is made of).  d_scratch: >= 4 bytes of device memory. */
int bn_b200_imad_peak_dev(uint32_t* d_scratch, uint32_t blocks, uint32_t iters, void* stream);

/* Pairing batches larger than `pairs` are processed in chunks of that many pairings so that the library-owned line buffer
 * (28 160 B per pairing) stays bounded; default 2^18 (7.4 GB), 0 restores the default. */
int bn_b200_set_max_chunk(size_t pairs);
/* A pairing call may run as `parts` sub-batches (of at least `min_pairs` pairings each) on as many internal streams,
 * forked from and joined to the caller's stream: results and stream ordering are unchanged, the kernels of one sub-batch
 * fill the idle block slots of another one's last wave.  parts = 0: the library decides per call
 * (default), 1: never, 2..4: always; min_pairs = 0: default (2048).  Same switch as the BN_B200_SPLIT environment variable. */
int bn_b200_set_split(int parts, size_t min_pairs);

/* Per-kernel device timing of the most recent pairing_batch[_dev] call, measured with CUDA events on the
 * stream the kernels were launched on (enable first; reading synchronises that stream).
 * ms[0] = line-schedule kernel, ms[1] = Miller-loop, inversion and final-exponentiation kernels (three launches). */
int bn_b200_set_profiling(int enable);
int bn_b200_last_pairing_kernel_ms(float ms[2]);
/* Same, three values: ms[0] = line schedule (k_pair_lines_duo), ms[1] = Miller loop incl. the preparation of the Fq12 inversion
 * (k_miller), ms[2] = batch-wide Fq inversion + final exponentiation (k_fq_inv_batch + k_fexp). */
int bn_b200_last_pairing_kernel_ms3(float ms[3]);
/* Name of the i-th kernel of the pairing path (i = 0, 1, 2: the three values of bn_b200_last_pairing_kernel_ms3), NULL beyond. */
const char* bn_b200_pairing_kernel_name(int i);
/* Library-owned device scratch per pairing in flight (line coefficients + flag), bytes. */
size_t bn_b200_scratch_bytes_per_pairing(void);
/* Number of kernels this library has launched since init (for the bench's gpu_launches claim). */
unsigned long long bn_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BN_B200_H */
