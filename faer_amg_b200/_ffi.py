"""ctypes binding of ``libfamg.so`` -- the same C ABI (``include/famg.h``) a Rust FFI shim binds.

There is no CPU fallback: if the CUDA extension is missing this module raises at import of the
library handle, and every compute entry point fails with ``FamgError`` when no sm_100 device is
usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FAMG_LIB", os.path.join(_HERE, "libfamg.so"))  # FAMG_LIB: A/B builds of the same ABI

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)
i64 = C.c_int64
f64 = C.c_double
cint = C.c_int


class CgInfoStruct(C.Structure):
    _fields_ = [("iter_count", C.c_int64), ("abs_residual", C.c_double), ("rel_residual", C.c_double)]


class FamgError(RuntimeError):
    """Non-zero ``famg_status``; the Rust shim would ``panic!`` with the same message."""

    def __init__(self, status: int, message: str):
        super().__init__(f"famg status {status}: {message}")
        self.status = status
        self.message = message


OK, ERR_INVALID, ERR_CUDA, ERR_ALLOC, ERR_NUMERIC, ERR_NO_CONVERGENCE, ERR_NOT_SPD, ERR_COMM, ERR_UNSUPPORTED = range(9)

# name -> argtypes; every function returns famg_status except the two string getters
SIGNATURES = {
    "famg_ctx_create": [cint, vpp],
    "famg_ctx_destroy": [vp],
    "famg_ctx_sync": [vp],
    "famg_ctx_stream": [vp, vpp],
    "famg_ctx_info": [vp, C.POINTER(cint), i64p, i64p, C.c_char_p, cint],
    "famg_ctx_launch_count": [vp, i64p],
    "famg_ctx_set_option": [vp, C.c_char_p, i64],
    "famg_ctx_trace_dump": [vp, C.c_char_p],
    "famg_set_num_threads": [cint],
    "famg_ctx_reserve": [vp, i64],
    "famg_csr_create": [vp, i64, i64, u64p, u64p, f64p, vpp],
    "famg_csr_create_from_triplets": [vp, i64, i64, i64, u64p, u64p, f64p, vpp],
    "famg_csr_retain": [vp],
    "famg_csr_destroy": [vp],
    "famg_csr_dims": [vp, i64p, i64p, i64p],
    "famg_csr_download": [vp, u64p, u64p, f64p],
    "famg_csr_plan": [vp, C.POINTER(cint), C.POINTER(cint), f64p, C.POINTER(cint)],
    "famg_csr_row_slab": [vp, i64, i64, vpp],
    "famg_gallery_g7": [vp, i64, i64, i64, vpp],
    "famg_gallery_g27": [vp, i64, i64, i64, f64, f64, vpp],
    "famg_vec_create": [vp, i64, i64, vpp],
    "famg_vec_destroy": [vp],
    "famg_vec_dims": [vp, i64p, i64p],
    "famg_vec_upload": [vp, f64p, i64],
    "famg_vec_download": [vp, f64p, i64],
    "famg_vec_fill": [vp, f64],
    "famg_vec_copy": [vp, vp],
    "famg_vec_axpby": [vp, f64, vp, f64],
    "famg_vec_ptr": [vp, vpp, i64p],
    "famg_vec_norm2": [vp, f64p],
    "famg_spmm": [vp, f64p, i64, f64p, i64, i64],
    "famg_spmm_dev": [vp, vp, vp],
    "famg_residual_dev": [vp, vp, vp, vp],
    "famg_spmm_add_dev": [vp, vp, vp],
    "famg_smoother_diag": [vp, cint, f64, vpp],
    "famg_smoother_diag_from_host": [vp, i64, f64p, vpp],
    "famg_smoother_cholesky": [vp, vpp],
    "famg_smoother_block": [vp, i64, u64p, u64p, vpp],
    "famg_smoother_block_vector": [vp, i64, i64, u64p, u64p, vpp],
    "famg_smoother_retain": [vp],
    "famg_smoother_destroy": [vp],
    "famg_smoother_dim": [vp, i64p],
    "famg_smoother_diag_download": [vp, f64p],
    "famg_smoother_apply": [vp, f64p, i64, f64p, i64, i64],
    "famg_smoother_apply_dev": [vp, vp, vp],
    "famg_smooth_dev": [vp, vp, vp, vp, cint],
    "famg_stationary_iteration_dev": [vp, vp, cint, vp],
    "famg_mg_create": [vp, vp, vpp],
    "famg_mg_add_level": [vp, vp, vp, vp, vp],
    "famg_mg_set_cycle": [vp, cint, cint],
    "famg_mg_levels": [vp, C.POINTER(cint)],
    "famg_mg_destroy": [vp],
    "famg_mg_apply": [vp, f64p, i64, f64p, i64, i64],
    "famg_mg_apply_dev": [vp, vp, vp],
    "famg_mg_cycle_bytes": [vp, i64, f64p],
    "famg_spgemm": [vp, vp, vpp],
    "famg_transpose": [vp, vpp],
    "famg_smooth_interpolation": [vp, vp, f64, vpp],
    "famg_galerkin": [vp, vp, cint, f64, vpp, vpp, vpp],
    "famg_galerkin_block": [vp, vp, i64, cint, f64, vpp, vpp, vpp],
    "famg_block_jacobi": [vp, i64, vp, vpp],
    "famg_smooth_p": [vp, vp, vp, vpp],
    "famg_tentative_p": [vp, i64, i64, i64, i64, f64p, i64, i64, u64p, u64p, vpp, f64p],
    "famg_thin_q": [i64, i64, f64p, i64],
    "famg_geometric_partition": [i64, i64, i64, i64, i64, i64, u64p, u64p, i64p],
    "famg_thin_q_dev": [vp],
    "famg_error_propagator_dev": [vp, vp, vp, vp],
    "famg_smooth_vector_dev": [vp, vp, i64, vp, f64p],
    "famg_vec_coldot": [vp, vp, f64p],
    "famg_smooth_vector_pc_dev": [vp, cint, vp, i64, vp, f64p],
    "famg_composite_create": [vp, vpp],
    "famg_composite_push": [vp, cint, vp],
    "famg_composite_len": [vp, i64p],
    "famg_composite_apply_dev": [vp, vp, vp],
    "famg_composite_destroy": [vp],
    "famg_strength_graph_create": [i64, u64p, u64p, f64p, i64, i64, f64p, i64, vpp],
    "famg_graph_create": [i64, u64p, u64p, f64p, vpp],
    "famg_graph_block_reduce": [vp, i64],
    "famg_graph_dims": [vp, i64p, i64p],
    "famg_graph_download": [vp, u64p, u64p, f64p],
    "famg_graph_destroy": [vp],
    "famg_partition_modularity": [vp, f64, f64, i64, u64p, i64p],
    "famg_pcg_solve": [vp, cint, vp, f64p, f64p, f64, f64, i64, cint, C.POINTER(CgInfoStruct)],
    "famg_pcg_solve_dev": [vp, cint, vp, vp, vp, f64, f64, i64, cint, C.POINTER(CgInfoStruct)],
    "famg_stationary_solve": [vp, cint, vp, f64p, f64p, f64, i64, i64p],
    "famg_comm_unique_id": [vp],
    "famg_comm_create": [vp, cint, cint, vp, vpp],
    "famg_comm_destroy": [vp],
    "famg_comm_allreduce_sum": [vp, f64p, cint],
    "famg_dist_mg_create": [vp, vp, C.POINTER(i64p), i64, vpp],
    "famg_dist_mg_destroy": [vp],
    "famg_dist_pcg_solve": [vp, f64p, f64p, f64, f64, i64, cint, C.POINTER(CgInfoStruct)],
    "famg_dist_pcg_solve_dev": [vp, vp, vp, f64, f64, i64, cint, C.POINTER(CgInfoStruct)],
    "famg_dist_spmv_dev": [vp, vp, vp],
    "famg_dist_mg_apply_dev": [vp, vp, vp],
    "famg_time_kernel": [vp, cint, cint, cint, C.POINTER(C.c_float)],
    "famg_gallery_g7_slab": [vp, i64, i64, i64, i64, i64, vpp],
    "famg_gallery_g27_slab": [vp, i64, i64, i64, f64, f64, i64, i64, vpp],
    "famg_comm_create_sim": [vp, cint, vpp],
    "famg_comm_dims": [vp, C.POINTER(cint), C.POINTER(cint), C.POINTER(cint)],
    "famg_comm_allgatherv_f64": [vp, C.POINTER(f64p), i64p, C.POINTER(f64p)],
    "famg_dmat_create": [vp, vpp, i64, i64p, vpp],
    "famg_dmat_finalize": [vp, cint],
    "famg_dmat_retain": [vp],
    "famg_dmat_destroy": [vp],
    "famg_dmat_info": [vp, i64p, i64p, i64p, i64p],
    "famg_dmat_local": [vp, cint, cint, vpp],
    "famg_dmat_gather": [vp, vpp],
    "famg_dist_coarsen": [vp, i64p, C.POINTER(u64p), C.POINTER(u64p), C.POINTER(f64p), cint, f64, vpp, vpp, vpp, C.POINTER(f64p)],
    "famg_dist_smooth_near_null": [vp, cint, C.POINTER(f64p)],
    "famg_dist_coarsen_dev": [vp, vpp, vpp, cint, f64, vpp, vpp, vpp, vpp],
    "famg_dist_smooth_near_null_dev": [vp, cint, vpp],
    "famg_partition_geometric_dev": [vp, i64, i64, i64, i64, i64, i64, vpp, i64p],
    "famg_partition_upload": [vp, i64, i64, u64p, u64p, vpp],
    "famg_partition_dims": [vp, i64p, i64p],
    "famg_partition_download": [vp, u64p, u64p],
    "famg_partition_destroy": [vp],
    "famg_tentative_p_dev": [vp, vp, vpp, vp],
    "famg_dist_mg_create_levels": [vp, cint, vpp, vpp, vpp, cint, f64, vp, vpp],
}

_lib = None


def lib():
    """Load ``libfamg.so`` (built in-tree by ``__graft_entry__.build()``). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()')."
            " faer_amg_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = cint
    L.famg_last_error.restype = C.c_char_p
    L.famg_last_error.argtypes = []
    L.famg_version.restype = C.c_char_p
    L.famg_version.argtypes = []
    _lib = L
    return L


def check(status: int):
    if status != OK:
        raise FamgError(status, lib().famg_last_error().decode("utf-8", "replace"))


def call(name: str, *args):
    check(getattr(lib(), name)(*args))
