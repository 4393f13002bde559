"""ctypes binding of include/nlo_cuda.h (libnlo_cuda.so).  Fails loudly when the library is
missing -- there is no Python / CPU fallback for the hot path."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLO_LIB", os.path.join(_PKG, "libnlo_cuda.so"))

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)
c_int_p = ctypes.POINTER(ctypes.c_int)


class SolveOptions(ctypes.Structure):
    _fields_ = [("max_iterations", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("parameter_tolerance", ctypes.c_double), ("gradient_tolerance", ctypes.c_double)]


class RegisterResult(ctypes.Structure):
    _fields_ = [("outer_iterations", ctypes.c_int32), ("inner_iterations", ctypes.c_int32),
                ("status", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("final_cost", ctypes.c_double), ("device_ms", ctypes.c_double),
                ("matched", ctypes.c_int64)]


class SolveResult(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("status", ctypes.c_int32),
                ("final_cost", ctypes.c_double), ("device_ms", ctypes.c_double)]


_VP = ctypes.c_void_p
# name -> (restype, argtypes); every symbol include/nlo_cuda.h declares
_SIGNATURES = {
    "nlo_abi_version": (ctypes.c_int, []),
    "nlo_visible_device_count": (ctypes.c_int, []),
    "nlo_context_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_VP)]),
    "nlo_context_create_multi": (ctypes.c_int, [c_int32_p, ctypes.c_int32, ctypes.POINTER(_VP)]),
    "nlo_context_device_count": (ctypes.c_int, [_VP]),
    "nlo_context_destroy": (ctypes.c_int, [_VP]),
    "nlo_last_error": (ctypes.c_char_p, [_VP]),
    "nlo_context_info": (ctypes.c_int, [_VP, c_int_p, c_int_p]),
    "nlo_synchronize": (ctypes.c_int, [_VP]),
    "nlo_set_loss": (ctypes.c_int, [_VP, ctypes.c_int, c_double_p]),
    "nlo_host_alloc": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_size_t]),
    "nlo_host_free": (ctypes.c_int, [_VP]),
    "nlo_debug_guard_report": (ctypes.c_int, [c_int32_p, c_int64_p, c_int64_p, c_int64_p]),
    "nlo_ndt_create": (ctypes.c_int, [_VP, ctypes.c_int64, ctypes.POINTER(_VP)]),
    "nlo_ndt_create_f32": (ctypes.c_int, [_VP, ctypes.c_int64, ctypes.POINTER(_VP)]),
    "nlo_ndt_create_batched": (ctypes.c_int, [_VP, ctypes.c_int32, c_int64_p, ctypes.POINTER(_VP)]),
    "nlo_ndt_upload": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, _VP, _VP, _VP]),
    "nlo_ndt_upload_f32": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, _VP, _VP, _VP]),
    "nlo_ndt_upload_aos": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, _VP, ctypes.c_size_t,
                                          ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                          ctypes.c_int]),
    "nlo_ndt_generate": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64,
                                        ctypes.c_double, c_double_p, c_double_p, c_double_p,
                                        c_int32_p, ctypes.c_double, c_double_p, c_double_p,
                                        c_uint8_p]),
    "nlo_ndt_generate_batched": (ctypes.c_int, [_VP, _VP, ctypes.c_uint64, ctypes.c_double,
                                                c_double_p, c_double_p, c_double_p, c_int32_p,
                                                ctypes.c_double, c_double_p, c_double_p, c_uint8_p]),
    "nlo_ndt_download": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_int64, c_double_p,
                                        c_double_p, c_double_p]),
    "nlo_ndt_download_problem": (ctypes.c_int, [_VP, _VP, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                                c_double_p, c_double_p, c_double_p]),
    "nlo_ingest_stats": (ctypes.c_int, [_VP, c_double_p, c_double_p]),
    "nlo_reproj_upload_aos": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, _VP, ctypes.c_size_t,
                                             ctypes.c_size_t, ctypes.c_size_t, c_double_p]),
    "nlo_reproj_create": (ctypes.c_int, [_VP, ctypes.c_int64, ctypes.POINTER(_VP)]),
    "nlo_reproj_create_batched": (ctypes.c_int, [_VP, ctypes.c_int32, c_int64_p, ctypes.POINTER(_VP)]),
    "nlo_reproj_upload": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, _VP, _VP, c_double_p]),
    "nlo_problem_destroy": (ctypes.c_int, [_VP, _VP]),
    "nlo_problem_size": (ctypes.c_int64, [_VP]),
    "nlo_ndt6_assemble": (ctypes.c_int, [_VP, _VP, ctypes.c_int32, c_double_p, ctypes.c_int64,
                                         ctypes.c_int64, c_double_p, c_double_p, c_double_p]),
    "nlo_ndt3_assemble": (ctypes.c_int, [_VP, _VP, ctypes.c_int32, c_double_p, ctypes.c_int64,
                                         ctypes.c_int64, c_double_p, c_double_p, c_double_p]),
    "nlo_reproj_assemble": (ctypes.c_int, [_VP, _VP, ctypes.c_int32, c_double_p, ctypes.c_int64,
                                           ctypes.c_int64, c_double_p, c_double_p, c_double_p]),
    "nlo_ndt6_solve": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(SolveOptions), c_double_p,
                                      ctypes.POINTER(SolveResult), c_double_p]),
    "nlo_ndt3_solve": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(SolveOptions), c_double_p,
                                      ctypes.POINTER(SolveResult), c_double_p]),
    "nlo_reproj_solve": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(SolveOptions), c_double_p,
                                        ctypes.POINTER(SolveResult), c_double_p]),
    "nlo_ndt6_solve_batched": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(SolveOptions), c_double_p,
                                              ctypes.POINTER(SolveResult)]),
    "nlo_ndt3_solve_batched": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(SolveOptions), c_double_p,
                                              ctypes.POINTER(SolveResult)]),
    "nlo_reproj_solve_batched": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(SolveOptions), c_double_p,
                                                ctypes.POINTER(SolveResult)]),
    "nlo_ndt_map_create": (ctypes.c_int, [_VP, c_double_p, c_int32_p, ctypes.c_double, c_double_p,
                                          c_double_p, c_uint8_p, ctypes.POINTER(_VP)]),
    "nlo_ndt_map_build": (ctypes.c_int, [_VP, ctypes.c_int64, _VP, ctypes.c_double, ctypes.c_int,
                                         ctypes.POINTER(_VP)]),
    "nlo_ndt_map_build_hashed": (ctypes.c_int, [_VP, ctypes.c_int64, _VP, ctypes.c_double, ctypes.c_int,
                                                ctypes.POINTER(_VP)]),
    "nlo_ndt_map_layout": (ctypes.c_int, [_VP, _VP, c_int32_p, c_int64_p]),
    "nlo_ndt_map_download_keys": (ctypes.c_int, [_VP, _VP, _VP]),
    "nlo_ndt_map_info": (ctypes.c_int, [_VP, _VP, c_double_p, c_int32_p, c_double_p, c_int64_p]),
    "nlo_ndt_map_download": (ctypes.c_int, [_VP, _VP, c_double_p, c_double_p, c_uint8_p]),
    "nlo_ndt_map_destroy": (ctypes.c_int, [_VP, _VP]),
    "nlo_scan_create": (ctypes.c_int, [_VP, ctypes.c_int64, _VP, ctypes.POINTER(_VP)]),
    "nlo_scan_destroy": (ctypes.c_int, [_VP, _VP]),
    "nlo_ndt_match": (ctypes.c_int, [_VP, _VP, _VP, c_double_p, ctypes.c_double, ctypes.c_int32,
                                     _VP, c_int64_p]),
    "nlo_ndt_register": (ctypes.c_int, [_VP, _VP, _VP, ctypes.POINTER(SolveOptions),
                                        ctypes.c_double, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int32, c_double_p, ctypes.POINTER(RegisterResult)]),
    "nlo_comm_unique_id": (ctypes.c_int, [_VP, c_uint8_p]),
    "nlo_comm_init_nccl": (ctypes.c_int, [_VP, c_uint8_p, ctypes.c_int32, ctypes.c_int32]),
    "nlo_comm_peer_export": (ctypes.c_int, [_VP, c_uint8_p]),
    "nlo_comm_peer_init": (ctypes.c_int, [_VP, c_uint8_p, ctypes.c_int32, ctypes.c_int32]),
    "nlo_problem_set_global_range": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_int64]),
    "nlo_comm_suspend": (ctypes.c_int, [_VP, ctypes.c_int32]),
    "nlo_comm_destroy": (ctypes.c_int, [_VP]),
}

_lib = None


def load():
    """Load libnlo_cuda.so and bind every declared symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libnlo_cuda.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`"
            " or nonlinear_optimizer_for_slam_b200/build.py); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def declared_symbols():
    return sorted(_SIGNATURES)
