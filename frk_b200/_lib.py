"""ctypes binding of libfrkb200.so (the C ABI declared in include/frkb200.h).

The product path has no CPU fallback: if the shared library is missing this module raises at
import time, and if no CUDA device is present ``frk_ctx_create`` fails with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrkb200.so")


class FRKError(RuntimeError):
    """Raised when a libfrkb200 entry point returns a non-zero status."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
            "frk_b200 has no CPU fallback and cannot run without its CUDA library.")
    return C.CDLL(LIB_PATH)


lib = _load()

c_i64 = C.c_int64
c_int = C.c_int
c_dbl = C.c_double
vp = C.c_void_p
p_i64 = C.POINTER(C.c_int64)
p_dbl = C.POINTER(C.c_double)
p_int = C.POINTER(C.c_int)

ALLREDUCE_FN = C.CFUNCTYPE(c_int, vp, vp, c_i64, c_int, c_int)
OBJECTIVE_FN = C.CFUNCTYPE(c_dbl, vp, c_int, p_dbl)

# name -> (restype, argtypes); must list every symbol include/frkb200.h declares
SIGNATURES = {
    "frk_abi_version": (c_int, []),
    "frk_last_error": (C.c_char_p, []),
    "frk_ctx_create": (c_int, [c_int, vp, C.POINTER(vp)]),
    "frk_ctx_destroy": (c_int, [vp]),
    "frk_sync": (c_int, [vp]),
    "frk_malloc": (c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "frk_free": (c_int, [vp, vp]),
    "frk_h2d": (c_int, [vp, vp, vp, C.c_size_t]),
    "frk_d2h": (c_int, [vp, vp, vp, C.c_size_t]),
    "frk_d2d": (c_int, [vp, vp, vp, C.c_size_t]),
    "frk_memset0": (c_int, [vp, vp, C.c_size_t]),
    "frk_launch_count": (c_i64, [vp]),
    "frk_use_graphs": (c_int, [vp, c_int]),
    "frk_use_deterministic": (c_int, [vp, c_int]),
    "frk_profile": (c_int, [vp, c_int]),
    "frk_profile_report": (c_int, [vp, C.c_char_p, C.c_size_t]),
    "frk_debug_clocks": (c_int, [vp]),
    "frk_comm_unique_id": (c_int, [vp]),
    "frk_comm_init": (c_int, [vp, c_int, c_int, vp]),
    "frk_comm_destroy": (c_int, [vp]),
    "frk_comm_rank": (c_int, [vp]),
    "frk_comm_world": (c_int, [vp]),
    "frk_comm_allreduce": (c_int, [vp, vp, c_i64, c_int]),
    "frk_comm_allreduce_host": (c_int, [vp, vp, c_i64, c_int]),
    "frk_comm_broadcast": (c_int, [vp, vp, c_i64, c_int]),
    "frk_basis_create": (c_int, [vp, c_int, c_dbl, c_int, c_int, c_int, vp, vp, vp, C.POINTER(vp)]),
    "frk_basis_destroy": (c_int, [vp]),
    "frk_basis_dist_matrix": (c_int, [vp, vp, c_int, vp, vp]),
    "frk_basis_nn_dist": (c_int, [vp, vp, c_int, vp]),
    "frk_csr_colsums": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp]),
    "frk_basis_n": (c_int, [vp]),
    "frk_basis_dim": (c_int, [vp]),
    "frk_eval_basis_count": (c_int, [vp, vp, c_i64, vp, c_i64, vp, p_i64]),
    "frk_eval_basis_fill": (c_int, [vp, vp, c_i64, vp, c_i64, vp, vp, vp, c_int]),
    "frk_csr_tensor_count": (c_int, [vp, c_i64, vp, vp, vp, p_i64]),
    "frk_csr_tensor_fill": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "frk_csr_row_normalise": (c_int, [vp, c_i64, vp, vp]),
    "frk_polygon_lookup_count": (c_int, [vp, vp, c_i64, vp, c_i64, vp, p_i64]),
    "frk_polygon_incidence_count": (c_int, [vp, vp, c_i64, vp, c_i64, vp, p_i64]),
    "frk_polygon_incidence_fill": (c_int, [vp, vp, c_i64, vp, c_i64, vp, c_i64, vp, vp, vp, vp, p_i64]),
    "frk_areal_incidence": (c_int, [vp, c_i64, vp, c_i64, vp, vp, vp, vp, p_i64, p_i64]),
    "frk_csr_aggregate_count": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, p_i64]),
    "frk_csr_aggregate_fill": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "frk_areal_aggregate": (c_int, [vp, c_i64, vp, vp, vp, c_int, c_int, vp, c_i64, vp]),
    "frk_st_bau_index": (c_int, [vp, c_i64, vp, vp, c_int, vp, c_i64, c_i64, c_i64, vp]),
    "frk_grid_lookup": (c_int, [vp, c_i64, vp, c_i64, vp, vp, vp, vp, vp]),
    "frk_grid_lookup_range": (c_int, [vp, c_i64, vp, c_i64, vp, vp, vp, vp, c_int, c_int, c_int, vp]),
    "frk_polygons_create": (c_int, [vp, c_int, vp, vp, vp, C.POINTER(vp)]),
    "frk_polygons_destroy": (c_int, [vp]),
    "frk_polygon_lookup": (c_int, [vp, vp, c_i64, vp, c_i64, c_int, c_int, vp]),
    "frk_bin_mean": (c_int, [vp, c_i64, vp, c_i64, c_int, vp, vp, vp, vp, p_i64]),
    "frk_bin_mean_sum": (c_int, [vp, c_i64, vp, c_i64, c_int, vp, C.c_uint, vp, vp, vp, p_i64]),
    "frk_grid_centroids": (c_int, [vp, c_i64, c_i64, c_int, vp, vp, vp]),
    "frk_csr_gather_count": (c_int, [vp, c_i64, vp, vp, vp, p_i64]),
    "frk_csr_gather_fill": (c_int, [vp, c_i64, vp, vp, vp, vp, vp, vp, vp]),
    "frk_gram_rhs": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, c_int, vp, vp, vp, c_dbl, vp]),
    "frk_csr_bucket_rows": (c_int, [vp, c_i64, vp, vp, c_int, vp, vp, vp, vp, vp]),
    "frk_rowbuckets_create": (c_int, [vp, c_i64, c_int, vp, vp, vp, C.POINTER(vp)]),
    "frk_rowbuckets_destroy": (c_int, [vp]),
    "frk_gram_rhs_bucketed": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, c_int, vp, vp, vp, c_dbl, vp]),
    "frk_quadform_spmv_bucketed": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, vp, vp]),
    "frk_quadform_spmv": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, vp]),
    "frk_mstep_sums": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, c_dbl, c_int, vp]),
    "frk_vec_stats": (c_int, [vp, c_i64, vp, c_dbl, vp]),
    "frk_square": (c_int, [vp, c_i64, vp, vp]),
    "frk_recip": (c_int, [vp, c_i64, vp, vp]),
    "frk_fill": (c_int, [vp, c_i64, c_dbl, vp]),
    "frk_scale": (c_int, [vp, c_i64, c_dbl, vp]),
    "frk_gather_rows": (c_int, [vp, c_i64, c_int, vp, vp, c_i64, vp, c_i64]),
    "frk_predict_finalize": (c_int, [vp, c_i64, c_int, c_dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "frk_scatter_add": (c_int, [vp, c_i64, vp, vp, vp]),
    "frk_resid": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp]),
    "frk_xalpha": (c_int, [vp, c_i64, c_int, vp, vp, vp]),
    "frk_potrf": (c_int, [vp, c_int, vp, p_dbl]),
    "frk_potri": (c_int, [vp, c_int, vp, vp, vp]),
    "frk_lower_mulv": (c_int, [vp, c_int, vp, vp, vp]),
    "frk_symv": (c_int, [vp, c_int, vp, vp, vp]),
    "frk_sym_from_upper_add": (c_int, [vp, c_int, vp, vp, vp]),
    "frk_scaled_identity": (c_int, [vp, c_int, c_dbl, vp]),
    "frk_add_diag": (c_int, [vp, c_int, vp, vp]),
    "frk_add_scaled_outer": (c_int, [vp, c_int, vp, vp, c_dbl, vp]),
    "frk_add_outer": (c_int, [vp, c_int, vp, vp, vp]),
    "frk_exp_kernel": (c_int, [vp, c_i64, vp, c_dbl, c_dbl, vp]),
    "frk_dot": (c_int, [vp, c_i64, vp, vp, p_dbl]),
    "frk_block_gather": (c_int, [vp, c_int, vp, c_int, vp, vp]),
    "frk_block_scatter": (c_int, [vp, c_int, vp, c_int, vp, vp]),
    "frk_block_gather_outer": (c_int, [vp, c_int, vp, vp, c_int, vp, vp]),
    "frk_kron_dot": (c_int, [vp, c_int, c_int, vp, vp, vp, p_dbl]),
    "frk_kron_scatter": (c_int, [vp, c_int, c_int, c_int, vp, vp, c_dbl, vp, vp]),
    "frk_dgemm_nt": (c_int, [vp, c_int, c_int, c_int, c_dbl, vp, c_int, vp, c_int, c_dbl, vp, c_int]),
    "frk_dgemm": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, c_dbl, vp, c_int, vp, c_int, c_dbl, vp, c_int]),
    "frk_model_create": (c_int, [vp, c_i64, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, ALLREDUCE_FN, vp, C.POINTER(vp)]),
    "frk_model_create_host": (c_int, [vp, c_i64, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "frk_model_destroy": (c_int, [vp]),
    "frk_model_predict_polygons": (c_int, [vp, c_i64, vp, vp, vp, vp, vp, vp, vp, c_int, c_i64, vp, vp, vp, c_i64, vp, vp, vp]),
    "frk_model_predict_areal": (c_int, [vp, c_i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "frk_model_set_lanes": (c_int, [vp, c_int]),
    "frk_model_set_vfs_csr": (c_int, [vp, vp, vp, vp]),
    "frk_polygon_lookup_host": (c_int, [vp, c_int, vp, vp, vp, c_i64, vp, vp]),
    "frk_model_create_implicit": (c_int, [vp, c_i64, c_int, vp, vp, c_i64, c_int, vp, vp, vp, vp, ALLREDUCE_FN, vp, C.POINTER(vp)]),
    "frk_rowbuckets_create_implicit": (c_int, [vp, c_i64, vp, vp, c_i64, c_int, C.POINTER(vp)]),
    "frk_model_coupled_groups": (c_int, [vp]),
    "frk_model_set_nm_distribute": (c_int, [vp, c_int]),
    "frk_model_set_parallel": (c_int, [vp, c_int, c_int, c_int]),
    "frk_model_set_blocks": (c_int, [vp, c_int, vp, C.POINTER(vp)]),
    "frk_model_set_blocks_basis": (c_int, [vp, c_int, vp, vp]),
    "frk_model_set_blocks_tensor": (c_int, [vp, c_int, c_int, c_int, vp, C.POINTER(vp), vp]),
    "frk_model_init": (c_int, [vp]),
    "frk_model_em_fit": (c_int, [vp, c_int, c_dbl, c_int, c_int, vp, vp, c_int, c_dbl, vp, p_int]),
    "frk_model_loglik": (c_int, [vp, p_dbl]),
    "frk_model_get": (c_int, [vp, C.c_char_p, vp]),
    "frk_model_set": (c_int, [vp, C.c_char_p, vp]),
    "frk_model_slot_device": (c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "frk_model_predict": (c_int, [vp, c_i64, vp, vp, vp, vp, vp, vp, vp, c_int, vp, vp, vp]),
    "frk_model_predict_host": (c_int, [vp, c_i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, c_int, c_i64, vp, vp, vp, vp, vp, vp]),
    "frk_optim_nelder_mead": (c_int, [OBJECTIVE_FN, vp, c_int, vp, c_int, c_dbl, vp, p_dbl, p_int, p_int]),
    "frk_optim_nelder_mead_speculative": (c_int, [OBJECTIVE_FN, vp, c_int, vp, c_int, c_dbl, c_int, vp, p_dbl, p_int, p_int, p_int]),
    "frk_optim_zeroin": (c_int, [OBJECTIVE_FN, vp, c_dbl, c_dbl, p_dbl]),
    "frk_eval_basis_host": (c_int, [vp, vp, c_i64, vp, c_int, vp, vp, vp, p_i64]),
    "frk_grid_lookup_host": (c_int, [vp, c_i64, vp, vp, vp, vp, vp, c_i64, vp]),
    "frk_quadform_host": (c_int, [vp, c_i64, c_int, vp, vp, vp, vp, vp, vp, vp]),
    "frk_chol2inv_host": (c_int, [vp, c_int, vp, vp, p_dbl]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)        # AttributeError here = the .so is stale; rebuild
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.frk_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status != 0:
        raise FRKError(last_error())


def call(name: str, *args):
    """Call an int-status entry point and raise FRKError with the library's message on failure."""
    check(getattr(lib, name)(*args))
