"""ctypes binding of include/cheq_b200.h -- the reference-side stub a maintainer would write in Rust
 restated for the Python test / bench harness.  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# CHEQ_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
_LIB_PATH = Path(os.environ.get("CHEQ_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libcheq_b200.so")

OK, NO_ATOMS, PARAMETER_NOT_FOUND, NOT_CONVERGED, LINALG, CUDA, INVALID_ARGUMENT, OUT_OF_MEMORY = range(8)


class Options(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("lambda_scale", C.c_double), ("damping_value", C.c_double),
                ("max_iterations", C.c_uint32), ("hydrogen_scf", C.c_uint8), ("basis", C.c_uint8),
                ("damping_kind", C.c_uint8), ("reserved", C.c_uint8)]


class External(C.Structure):
    _fields_ = [("num_point_charges", C.c_uint64), ("xyz", C.c_void_p), ("charge", C.c_void_p),
                ("radius", C.c_void_p), ("n", C.c_void_p), ("field", C.c_double * 3)]


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_host_prepare", C.c_double), ("ms_h2d", C.c_double),
                ("ms_assemble", C.c_double), ("ms_factor_once", C.c_double), ("ms_scf", C.c_double),
                ("ms_d2h", C.c_double), ("kernel_launches", C.c_uint64), ("n_atoms", C.c_uint64),
                ("n_hydrogen", C.c_uint64), ("n_padded", C.c_uint64), ("iterations", C.c_uint32),
                ("pair_types", C.c_uint32), ("flops_factor", C.c_double), ("bytes_assemble", C.c_double),
                ("lu_fallback_frames", C.c_uint64), ("gemm_ms", C.c_double), ("gemm_flops", C.c_double),
                ("gemm_launches", C.c_uint64), ("potrf_ms", C.c_double), ("backsolve_ms", C.c_double),
                ("krylov_iterations", C.c_uint64), ("krylov_steps", C.c_uint64), ("krylov_ms", C.c_double),
                ("krylov_bytes", C.c_double), ("matfree_matvecs", C.c_uint64), ("matfree_pairs", C.c_double),
                ("tc_gemm_launches", C.c_uint64), ("tc_gemm_flops", C.c_double), ("tc_gemm_ops", C.c_double),
                ("tc_gemm_ms", C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/cheq_b200.h declares: (restype, argtypes)
_P, _D, _U64, _I32, _U8 = C.c_void_p, C.c_double, C.c_uint64, C.c_int32, C.c_uint8
SIGNATURES = {
    "cheq_ctx_create": (_I32, [_I32, C.POINTER(_P)]),
    "cheq_ctx_destroy": (None, [_P]),
    "cheq_last_error": (C.c_char_p, [_P]),
    "cheq_options_default": (None, [C.POINTER(Options)]),
    "cheq_get_stats": (_I32, [_P, C.POINTER(Stats)]),
    "cheq_ctx_stream": (_P, [_P]),
    "cheq_set_profiling": (_I32, [_P, _I32]),
    "cheq_get_scf_trace": (_I32, [_P, C.c_uint32, _P, _P, _P, C.POINTER(C.c_uint32)]),
    "cheq_set_krylov": (_I32, [_P, _I32, _I32, _D, _I32]),
    "cheq_set_ozaki": (_I32, [_P, _I32, _I32, _I32]),
    "cheq_dist_unique_id": (_I32, [_P]),
    "cheq_ctx_dist_init": (_I32, [_P, _I32, _I32, _P]),
    "cheq_ctx_dist_info": (_I32, [_P, C.POINTER(_I32), C.POINTER(_I32)]),
    "cheq_set_factor_mode": (_I32, [_P, _I32, _I32]),
    "cheq_dist_plan": (_I32, [_I32, _I32, _I32, _I32, _P, _P]),
    "cheq_solve": (_I32, [_P, _U64, _P, _P, _P, _P, _P, _D, C.POINTER(Options), C.POINTER(External), _P,
                          C.POINTER(_D), C.POINTER(C.c_uint32), C.POINTER(_D)]),
    "cheq_solve_iterative": (_I32, [_P, _U64, _P, _P, _P, _P, _P, _D, C.POINTER(Options), C.POINTER(External), _P,
                                    C.POINTER(_D), C.POINTER(C.c_uint32), C.POINTER(_D)]),
    "cheq_solve_device": (_I32, [_P, _U64, _P, _P, _P, _P, _P, _D, C.POINTER(Options), C.POINTER(External), _P,
                                 C.POINTER(_D), C.POINTER(C.c_uint32), C.POINTER(_D)]),
    "cheq_solve_batch": (_I32, [_P, _U64, _U64, _P, _P, _P, _P, _P, _D, C.POINTER(Options), _P, _P, _P, _P, _P]),
    "cheq_pair_integrals": (_I32, [_P, _U64, _P, _P, _P, _P, _P, _D, _U8, _P]),
    "cheq_assemble_matrix": (_I32, [_P, _U64, _P, _P, _P, _P, _D, _U8, _P]),
    "cheq_external_potential": (_I32, [_P, _U64, _P, _P, _P, C.POINTER(External), _D, _U8, _P]),
    "cheq_dbg_cholesky_solve": (_I32, [_P, _U64, _P, _U64, _P, _P, _P]),
    "cheq_dbg_gemm_nt": (_I32, [_P, _U64, _U64, _U64, _P, _P, _P, _I32]),
    "cheq_dbg_gemm_ozaki": (_I32, [_P, _U64, _U64, _U64, _P, _P, _P, _I32, _I32, _I32, _P]),
    "cheq_dbg_potrf_cycles": (_I32, [_P, _P, _I32]),
    "cheq_dbg_dfma_rate": (_I32, [_P, C.POINTER(_D)]),
    "cheq_dbg_gemm_bench": (_I32, [_P, _U64, _U64, _U64, _I32, _I32, _I32, C.POINTER(_D)]),
    "cheq_host_sto_table_eval": (_D, [_I32, _D, _I32, _D, _D, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32),
                                      C.POINTER(_D)]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def lib():
    """Loads the CUDA extension.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m cheq_b200.build` "
                              "(the QEq engine has no CPU fallback)")
        handle = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
