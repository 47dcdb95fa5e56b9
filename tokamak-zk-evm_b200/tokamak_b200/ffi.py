"""ctypes binding of libtokamak_b200.so (include/tokamak_b200.h).

The CUDA library is the only implementation: if the shared object is missing or no CUDA device is
usable this module raises -- there is no CPU fallback (and nothing here imports oracle/).
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libtokamak_b200.so")

c_void_p, c_size_t, c_int32, c_int64, c_uint32, c_uint64 = (
    ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64)
P = ctypes.POINTER

# every exported symbol of include/tokamak_b200.h with its argument types (return is int32 unless noted)
SIGNATURES = {
    "tkm_ctx_create": [c_int32, P(c_void_p)],
    "tkm_ctx_destroy": [c_void_p],
    "tkm_ctx_set_stream": [c_void_p, c_void_p],
    "tkm_ctx_sync": [c_void_p],
    "tkm_dev_alloc": [c_void_p, c_size_t, P(c_void_p)],
    "tkm_dev_free": [c_void_p, c_void_p],
    "tkm_memcpy_h2d": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_memcpy_d2h": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_ntt_domain_init": [c_void_p, c_uint32],
    "tkm_ntt_domain_release": [c_void_p],
    "tkm_ntt_domain_log2": [c_void_p, P(c_int32)],
    "tkm_root_of_unity": [c_uint32, c_void_p],
    "tkm_fr_to_mont": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_from_mont": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_vec_op": [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_vec_scale": [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_vec_inv": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_vec_op_host": [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_vec_fill": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_mul_x_minus_one": [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t],
    "tkm_fr_transpose": [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t],
    "tkm_fr_vec_reduce": [c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p],
    "tkm_fr_outer_product": [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t],
    "tkm_fr_powers": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_fr_gather": [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p],
    "tkm_fr_scatter_from_table": [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t],
    "tkm_fr_suffix_product": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_bintt": [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_int32, c_void_p, c_void_p],
    "tkm_bintt_host": [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_int32, c_void_p, c_void_p],
    "tkm_ntt_batch": [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_int32, c_int32, c_void_p],
    "tkm_msm_g1_host": [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "tkm_msm_g1": [c_void_p, c_void_p, c_int32, c_void_p, c_size_t, c_void_p],
    "tkm_ntt_batch_scatter": [c_void_p, c_void_p, c_size_t, c_size_t, c_int32, c_int32, c_void_p, c_void_p, c_uint32, c_uint64, c_uint64, c_uint64],
    "tkm_g1_bases_to_mont": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_g1_bases_from_mont": [c_void_p, c_void_p, c_void_p, c_size_t],
    "tkm_msm_g1_rect": [c_void_p, c_void_p, c_int32, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p],
    "tkm_msm_g1_indexed": [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p],
    "tkm_g1_fixed_base_mul": [c_void_p, c_void_p, c_void_p, c_int32, c_size_t, c_void_p],
    "tkm_g1_add": [c_void_p, c_void_p, c_void_p, c_void_p],
    "tkm_g1_mul": [c_void_p, c_void_p, c_void_p, c_void_p],
    "tkm_g1_sum": [c_void_p, c_void_p, c_size_t, c_void_p],
    "tkm_crs_upload": [c_void_p, c_void_p, c_size_t, c_size_t, P(c_void_p)],
    "tkm_crs_upload_mont": [c_void_p, c_void_p, c_size_t, c_size_t, P(c_void_p)],
    "tkm_crs_from_device": [c_void_p, c_void_p, c_size_t, c_size_t, c_int32, P(c_void_p)],
    "tkm_crs_precompute": [c_void_p, c_void_p, c_uint32],
    "tkm_crs_free": [c_void_p, c_void_p],
    "tkm_crs_device_ptr": [c_void_p, P(c_void_p), P(c_size_t), P(c_size_t)],
    "tkm_poly_from_coeffs_host": [c_void_p, c_void_p, c_size_t, c_size_t, P(c_void_p)],
    "tkm_poly_from_evals_host": [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_from_device": [c_void_p, c_void_p, c_size_t, c_size_t, P(c_void_p)],
    "tkm_poly_zero": [c_void_p, c_size_t, c_size_t, P(c_void_p)],
    "tkm_r1cs_uvw_polys": [c_void_p, ctypes.c_uint32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                           c_void_p, c_size_t, c_size_t, c_size_t, P(c_void_p), P(c_void_p), P(c_void_p)],
    "tkm_poly_clone": [c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_free": [c_void_p, c_void_p],
    "tkm_poly_shape": [c_void_p, P(c_size_t), P(c_size_t)],
    "tkm_poly_device_ptr": [c_void_p, P(c_void_p)],
    "tkm_poly_copy_coeffs_host": [c_void_p, c_void_p, c_void_p],
    "tkm_poly_to_evals_host": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "tkm_poly_ntt_inplace": [c_void_p, c_void_p, c_int32, c_void_p, c_void_p],
    "tkm_poly_find_degree": [c_void_p, c_void_p, P(c_int64), P(c_int64)],
    "tkm_poly_resize": [c_void_p, c_void_p, c_size_t, c_size_t],
    "tkm_poly_optimize_size": [c_void_p, c_void_p],
    "tkm_poly_mul_monomial": [c_void_p, c_void_p, c_size_t, c_size_t, P(c_void_p)],
    "tkm_poly_axpby": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_lincomb": [c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, P(c_void_p)],
    "tkm_polyexpr_eval": [c_void_p, c_void_p, c_uint32, c_void_p, c_uint32, c_void_p, c_uint32, c_size_t, c_size_t, P(c_void_p)],
    "tkm_msm_tree_stats": [c_void_p, P(c_uint32), c_void_p],
    "tkm_poly_kernel_time_last": [c_void_p, P(ctypes.c_float)],
    "tkm_poly_add_scalar": [c_void_p, c_void_p, c_void_p],
    "tkm_poly_mul": [c_void_p, c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_scale_coeffs": [c_void_p, c_void_p, c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_eval": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "tkm_poly_eval_x": [c_void_p, c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_eval_y": [c_void_p, c_void_p, c_void_p, P(c_void_p)],
    "tkm_poly_div_by_vanishing": [c_void_p, c_void_p, c_size_t, c_size_t, P(c_void_p), P(c_void_p)],
    "tkm_poly_div_by_ruffini": [c_void_p, c_void_p, c_void_p, c_void_p, P(c_void_p), P(c_void_p), c_void_p],
    "tkm_poly_divide_uni": [c_void_p, c_void_p, c_void_p, c_int32, P(c_void_p), P(c_void_p)],
    "tkm_poly_commit": [c_void_p, c_void_p, c_void_p, c_void_p],
    "tkm_poly_commit_begin": [c_void_p, c_void_p, c_void_p, P(c_int32)],
    "tkm_commit_end": [c_void_p, c_int32, c_void_p],
    "tkm_msm_g1_begin": [c_void_p, c_void_p, c_int32, c_void_p, c_size_t, P(c_int32)],
    "tkm_msm_g1_indexed_begin": [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, P(c_int32)],
    "tkm_comm_unique_id": [c_void_p],
    "tkm_comm_init": [c_void_p, c_void_p, c_int32, c_int32],
    "tkm_comm_destroy": [c_void_p],
    "tkm_comm_rank": [c_void_p, P(c_int32), P(c_int32)],
    "tkm_msm_g1_sharded": [c_void_p, c_void_p, c_int32, c_void_p, c_size_t, c_void_p],
    "tkm_bintt_sharded": [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_int32, c_void_p, c_void_p],
    "tkm_host_parse_hex_scalars": [ctypes.c_char_p, c_size_t, c_void_p, c_size_t, P(c_size_t)],
    "tkm_host_parse_r1cs": [ctypes.c_char_p, c_size_t, P(c_uint32), P(c_uint32), P(c_size_t), c_void_p, c_void_p, c_void_p],
    "tkm_event_time_begin": [c_void_p],
    "tkm_event_time_end": [c_void_p, P(ctypes.c_float)],
    "tkm_launch_count": [c_void_p, P(c_uint64)],
    "tkm_host_keccak256": [ctypes.c_char_p, c_size_t, ctypes.c_char_p],
    "tkm_microbench": [c_void_p, c_int32, P(ctypes.c_double)],
    "tkm_kernel_time_last": [c_void_p, P(ctypes.c_float)],
}
STRING_FUNCS = ("tkm_last_error", "tkm_version")


class TkmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"tkm status {status}: {message}")
        self.status = status


_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C tokamak-zk-evm_b200` (or __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int32
    for name in STRING_FUNCS:
        getattr(lib, name).restype = ctypes.c_char_p
        getattr(lib, name).argtypes = []
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise TkmError(status, load().tkm_last_error().decode())
