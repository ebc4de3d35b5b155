"""ctypes loader for libbn_b200.so.  Fails loudly: no library or no GPU => exception, never a CPU path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BN_B200_SO") or os.path.join(_HERE, "libbn_b200.so")  # env override: A/B builds only

EXPORTS = [
    "bn_b200_alloc_pinned", "bn_b200_device_count", "bn_b200_device_error", "bn_b200_fq_mul_chain",
    "bn_b200_fq_mul_chain_dev", "bn_b200_fq_sqr_chain", "bn_b200_fq_sqr_chain_dev", "bn_b200_fr_decode_batch",
    "bn_b200_fr_decode_batch_dev", "bn_b200_fr_encode_batch", "bn_b200_fr_encode_batch_dev", "bn_b200_fr_one",
    "bn_b200_fr_op_batch", "bn_b200_fr_op_batch_dev", "bn_b200_fr_pow_batch", "bn_b200_fr_pow_batch_dev",
    "bn_b200_free_pinned", "bn_b200_g1_add_batch", "bn_b200_g1_add_batch_dev", "bn_b200_g1_check_batch",
    "bn_b200_g1_check_batch_dev", "bn_b200_g1_decode_batch", "bn_b200_g1_decode_batch_dev",
    "bn_b200_g1_double_batch", "bn_b200_g1_double_batch_dev", "bn_b200_g1_encode_batch",
    "bn_b200_g1_encode_batch_dev", "bn_b200_g1_eq_batch", "bn_b200_g1_eq_batch_dev", "bn_b200_g1_mul_batch",
    "bn_b200_g1_mul_batch_dev", "bn_b200_g1_neg_batch", "bn_b200_g1_neg_batch_dev", "bn_b200_g1_normalize_batch",
    "bn_b200_g1_normalize_batch_dev", "bn_b200_g1_one", "bn_b200_g1_op_batch", "bn_b200_g1_op_batch_dev",
    "bn_b200_g1_sub_batch", "bn_b200_g1_sub_batch_dev", "bn_b200_g1_zero", "bn_b200_g2_add_batch",
    "bn_b200_g2_add_batch_dev", "bn_b200_g2_check_batch", "bn_b200_g2_check_batch_dev", "bn_b200_g2_decode_batch",
    "bn_b200_g2_decode_batch_dev", "bn_b200_g2_double_batch", "bn_b200_g2_double_batch_dev",
    "bn_b200_g2_encode_batch", "bn_b200_g2_encode_batch_dev", "bn_b200_g2_eq_batch", "bn_b200_g2_eq_batch_dev",
    "bn_b200_g2_mul_batch", "bn_b200_g2_mul_batch_dev", "bn_b200_g2_neg_batch", "bn_b200_g2_neg_batch_dev",
    "bn_b200_g2_normalize_batch", "bn_b200_g2_normalize_batch_dev", "bn_b200_g2_one", "bn_b200_g2_op_batch",
    "bn_b200_g2_op_batch_dev", "bn_b200_g2_sub_batch", "bn_b200_g2_sub_batch_dev", "bn_b200_g2_zero",
    "bn_b200_gt_exp_by_neg_z_batch", "bn_b200_gt_exp_by_neg_z_batch_dev", "bn_b200_gt_inv_batch",
    "bn_b200_gt_inv_batch_dev", "bn_b200_gt_mul_batch", "bn_b200_gt_mul_batch_dev", "bn_b200_gt_one",
    "bn_b200_gt_pow_batch", "bn_b200_gt_pow_batch_dev", "bn_b200_imad_peak_dev", "bn_b200_init",
    "bn_b200_init_multi", "bn_b200_last_error", "bn_b200_last_pairing_kernel_ms", "bn_b200_last_pairing_kernel_ms3",
    "bn_b200_launch_count", "bn_b200_num_lines", "bn_b200_pairing_batch", "bn_b200_pairing_batch_dev",
    "bn_b200_pairing_batch_gather_dev", "bn_b200_pairing_kernel_name", "bn_b200_pairing_pow_batch",
    "bn_b200_pairing_pow_batch_dev", "bn_b200_scratch_bytes_per_pairing", "bn_b200_set_max_chunk",
    "bn_b200_set_min_shard", "bn_b200_set_profiling", "bn_b200_set_split", "bn_b200_shutdown", "bn_b200_sm_count",
]


class BnB200Error(RuntimeError):
    pass


_lib = None


def load():
    """dlopen the library (does not touch the GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BnB200Error(
                "%s is missing: build it with `python -m bn_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        lib.bn_b200_last_error.restype = ctypes.c_char_p
        lib.bn_b200_launch_count.restype = ctypes.c_ulonglong
        lib.bn_b200_pairing_kernel_name.restype = ctypes.c_char_p
        lib.bn_b200_scratch_bytes_per_pairing.restype = ctypes.c_size_t
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise BnB200Error("bn_b200 error %d: %s" % (rc, load().bn_b200_last_error().decode()))


_inited = None


def init(device: int = 0):
    """Bind this process to one CUDA device.  Raises if no device / not sm_100-class."""
    global _inited
    if _inited != ("one", device) or load().bn_b200_device_count() == 0:
        check(load().bn_b200_init(int(device)))
        _inited = ("one", device)
    return load()


def init_multi(n_gpus: int = 0):
    """Bind devices 0..n_gpus-1 (0: all visible) to this process; host-pointer batches are then sharded over them."""
    global _inited
    if _inited != ("multi", n_gpus) or load().bn_b200_device_count() == 0:
        check(load().bn_b200_init_multi(int(n_gpus)))
        _inited = ("multi", n_gpus)
    return load()


def shutdown():
    """Release every bound device (streams, scratch, staging)."""
    global _inited
    check(load().bn_b200_shutdown())
    _inited = None


def ensure(device=None):
    """The binding the batch helpers use: an explicit device index rebinds to that one device; None keeps whatever is
    bound (bn_b200.init / init_multi), binding device 0 if nothing is."""
    if device is not None:
        return init(device)
    if _inited is None or load().bn_b200_device_count() == 0:
        return init(0)
    return load()


def pairing_kernel_names(lib=None):
    lib = lib or load()
    out = []
    for i in range(3):
        s = lib.bn_b200_pairing_kernel_name(i)
        out.append(s.decode() if s else "kernel%d" % i)
    return out
