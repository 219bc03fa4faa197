"""ctypes binding of libdinomc.so (the C ABI declared in include/dinomc.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError
is raised.  Build it with `./build.sh` (or `python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMC_LIB") or os.path.join(_HERE, "libdinomc.so")   # DMC_LIB: debug builds only (tools/gemm_trace.py)

DMC_F32, DMC_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_NORMALIZE_BWD, ACT_GELU_DG, ACT_MUL_AUX = 0, 1, 2, 3, 4, 5

i32, i64, f32, vp, sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class GemmArgs(C.Structure):
    """Mirror of `struct dmc_gemm_args` (include/dinomc.h)."""
    _fields_ = [
        ("M", i64), ("N", i64), ("K", i64),
        ("A", vp), ("lda", i64), ("a_mn_major", i32),
        ("B", vp), ("ldb", i64), ("b_mn_major", i32),
        ("A_lo", vp), ("B_lo", vp),
        ("in_dtype", i32),
        ("D", vp), ("ldd", i64), ("out_dtype", i32),
        ("col_scale", vp), ("bias", vp), ("alpha_dev", vp), ("alpha", f32),
        ("act", i32),
        ("aux", vp), ("ldaux", i64), ("aux_dtype", i32),
        ("split_k", i32),
        ("workspace", vp), ("workspace_bytes", sz),
        ("max_ctas", i32),
        ("stat_scale", f32), ("stat_center", vp), ("stat_row_partials", vp), ("stat_colsum_partials", vp),
        ("stat_bound", vp),
        ("stat_bound2", vp),
        ("row_scale", vp), ("row_eps", f32),
    ]


# name -> (restype, argtypes): every symbol include/dinomc.h declares.
SIGNATURES = {
    "dmc_version": (C.c_int, []),
    "dmc_last_error_string": (C.c_char_p, []),
    "dmc_device_check": (C.c_int, [C.c_int]),
    "dmc_set_pdl": (C.c_int, [C.c_int]),
    "dmc_set_streaming_ctas": (C.c_int, [C.c_int]),
    "dmc_gemm_workspace_bytes": (sz, [i64, i64, i64, i32]),
    "dmc_gemm_stats_parts": (i64, [i64]),
    "dmc_gemm": (C.c_int, [C.POINTER(GemmArgs), vp]),
    "dmc_gemm_simt": (C.c_int, [C.POINTER(GemmArgs), vp]),
    "dmc_split_tf32": (C.c_int, [vp, vp, vp, i64, vp]),
    "dmc_cast_f32_to_bf16": (C.c_int, [vp, vp, i64, vp]),
    "dmc_cast_f32_to_bf16_batch": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, vp]),
    "dmc_cast_bf16_to_f32_batch": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, vp]),
    "dmc_colsum_workspace_bytes": (sz, [i64, i64]),
    "dmc_colsum": (C.c_int, [vp, i32, i64, i64, i64, vp, vp, sz, vp]),
    "dmc_normalize_rows_fwd": (C.c_int, [vp, i32, i64, i64, i64, f32, vp, vp, vp, vp, vp]),
    "dmc_normalize_rows_bwd": (C.c_int, [vp, vp, vp, i64, i64, f32, vp, i32, vp]),
    "dmc_weightnorm_fwd": (C.c_int, [vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp]),
    "dmc_weightnorm_bwd": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp, vp]),
    "dmc_weightnorm_bwd_bf16": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp, vp]),
    "dmc_teacher_workspace_bytes": (sz, [i64, i64]),
    "dmc_teacher_stats_colsum": (C.c_int, [vp, i32, i64, i64, i64, vp, f32, vp, vp, vp, sz, vp]),
    "dmc_teacher_stats_colsum_bounded": (C.c_int, [vp, i32, i64, i64, i64, vp, f32, vp, vp, vp, vp, sz, vp]),
    "dmc_absmax": (C.c_int, [vp, i64, vp, vp]),
    "dmc_teacher_finalize": (C.c_int, [vp, vp, i64, i64, i64, i64, vp, vp, vp]),
    "dmc_rowdot": (C.c_int, [vp, i32, i64, i64, vp, vp, vp]),
    "dmc_lse_finalize": (C.c_int, [vp, i64, i64, vp, vp]),
    "dmc_ce_fused": (C.c_int, [vp, i32, i64, vp, i32, i64, vp, vp, vp, i64, i32, i32, i64, f32, f32, vp, i32, i64, vp, vp, sz, vp]),
    "dmc_scale_inplace_if": (C.c_int, [vp, i32, i64, vp, f32, vp]),
    "dmc_center_update": (C.c_int, [vp, vp, vp, i64, f32, f32, f32, vp]),
    "dmc_ce_workspace_bytes": (sz, [i64, i32, i32, i64]),
    "dmc_ce_fwd": (C.c_int, [vp, i32, i64, vp, i32, i64, vp, vp, i64, i32, i32, i64, f32, f32, vp, vp, vp, sz, vp]),
    "dmc_ce_bwd": (C.c_int, [vp, i32, i64, vp, i32, i64, vp, vp, vp, vp, i64, i32, i32, i64, f32, f32, vp, i32, i64, vp]),
    "dmc_ema_plan_bytes": (sz, [C.POINTER(i64), i64]),
    "dmc_ema_build_plan": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i64, vp, sz, C.POINTER(i64)]),
    "dmc_ema_multi_tensor": (C.c_int, [vp, i64, f32, f32, vp]),
    "dmc_ema_plan2_bytes": (sz, [C.POINTER(i64), i64, i64, i64]),
    "dmc_ema_build_plan2": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), i64, i64, i64, i64, vp, vp, vp,
                                      vp, sz, C.POINTER(i64)]),
    "dmc_ema_multi_tensor2": (C.c_int, [vp, i64, vp, vp]),
    "dmc_xrank_signal_bytes": (sz, [i32, i32]),
    "dmc_xrank_allreduce": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), i64, i32, i32, i32, f32, i32, i32, C.POINTER(vp),
                                      C.POINTER(i64), C.POINTER(i64), vp]),
    "dmc_clip_plan_bytes": (sz, [C.POINTER(i64), i64]),
    "dmc_clip_build_plan": (C.c_int, [C.POINTER(vp), C.POINTER(i64), i64, vp, sz, C.POINTER(i64)]),
    "dmc_clip_grads": (C.c_int, [vp, i64, f32, vp, vp, sz, vp]),
    "dmc_adamw_plan_bytes": (sz, [C.POINTER(i64), i64]),
    "dmc_adamw_build_plan": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i64, vp, sz, C.POINTER(i64)]),
    "dmc_adamw_multi_tensor": (C.c_int, [vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, i64, vp]),
    "dmc_lars_plan_bytes": (sz, [C.POINTER(i64), i64]),
    "dmc_lars_build_plan": (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), i64, vp, sz, C.POINTER(i64)]),
    "dmc_lars_multi_tensor": (C.c_int, [vp, i64, C.c_double, C.c_double, C.c_double, C.c_double, vp, sz, vp]),
}

_lib = None


def load():
    """Load libdinomc.so once; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run ./build.sh at the repo root. "
            "dinomc_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dmc_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libdinomc {what} failed with code {rc}: {msg}")
