"""ctypes binding of ``liboptconpy_b200.so`` (the C ABI declared in
``include/optconpy_b200.h``).  There is no fallback: if the library is missing
or fails to load, importing the compute modules raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'liboptconpy_b200.so')

i64, i32p, f64p, vp = C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p

# name -> (restype, argtypes); pointers are passed as integers (void*)
PROTOTYPES = {
    'ocb_last_error': (C.c_char_p, []),
    'ocb_version': (C.c_int, []),
    'ocb_launch_count': (i64, []),
    'ocb_spmm': (C.c_int, [i64, i64, i32p, i32p, f64p, f64p, i64, f64p, i64, i64,
                           C.c_double, C.c_double, vp]),
    'ocb_lu_create': (C.c_int, [C.POINTER(vp), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'ocb_lu_destroy': (C.c_int, [vp]),
    'ocb_lu_pack_host': (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64,
                                   C.POINTER(vp), C.POINTER(i64)]),
    'ocb_lu_pack_host_into': (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, i64,
                                        C.POINTER(i64)]),
    'ocb_lu_pack_host_checked': (C.c_int, [i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, i64,
                                           C.POINTER(vp), C.POINTER(i64), vp, vp, vp,
                                           C.POINTER(C.c_double)]),
    'ocb_host_free': (None, [vp]),
    'ocb_set_sync_mode': (C.c_int, [C.c_int]),
    'ocb_lu_create_from_image': (C.c_int, [C.POINTER(vp), vp, i64, vp, vp]),
    'ocb_lu_info': (C.c_int, [vp, C.POINTER(i64)]),
    'ocb_lu_stats': (C.c_int, [vp, C.POINTER(i64)]),
    'ocb_lu_panel_levels': (i64, [vp, vp, i64]),
    'ocb_debug_trace': (C.c_int, [vp, i64]),
    'ocb_lu_program_create': (C.c_int, [C.POINTER(vp), i64, vp, vp, vp, vp, vp, vp, i64]),
    'ocb_lu_program_destroy': (C.c_int, [vp]),
    'ocb_lu_program_template_hits': (i64, []),
    'ocb_lu_program_info': (C.c_int, [vp, C.POINTER(i64)]),
    'ocb_lu_program_solve_host': (C.c_int, [vp, vp, vp, vp, vp, i64, C.c_double, C.POINTER(i64)]),
    'ocb_lu_program_export': (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    'ocb_order_nd': (C.c_int, [i64, vp, vp, i64, vp]),
    'ocb_order_delay_zero_diagonals': (C.c_int, [i64, vp, vp, vp, vp, vp]),
    'ocb_refactor_create': (C.c_int, [C.POINTER(vp), i64, vp, vp, vp, vp]),
    'ocb_refactor_destroy': (C.c_int, [vp]),
    'ocb_refactor_info': (C.c_int, [vp, C.POINTER(i64)]),
    'ocb_refactor_structure': (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    'ocb_refactor_numeric': (C.c_int, [vp, vp, vp, vp]),
    'ocb_lu_solve_ws_bytes': (i64, [vp, i64]),
    'ocb_lu_solve': (C.c_int, [vp, f64p, i64, i64, f64p, i64, i64, i64, vp, i64, vp]),
    'ocb_prof_enable': (C.c_int, [C.c_int]),
    'ocb_prof_collect': (C.c_int, [C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]),
    'ocb_gram_ws_bytes': (i64, [i64, i64, i64]),
    'ocb_gram': (C.c_int, [f64p, i64, i64, f64p, i64, i64, i64, f64p, i64, vp, i64, vp]),
    'ocb_fp64_peak': (C.c_int, [C.c_int, i64, i64, C.POINTER(C.c_double), vp, vp]),
    'ocb_tall_gemm': (C.c_int, [f64p, i64, i64, i64, f64p, i64, i64, f64p, i64,
                                C.c_double, C.c_double, vp]),
    'ocb_sym_eig': (C.c_int, [f64p, i64, i64, f64p, f64p, i64, C.POINTER(C.c_int32), vp]),
    'ocb_compress_ws_bytes': (i64, [i64, i64, i64]),
    'ocb_compress': (C.c_int, [f64p, i64, i64, i64, C.c_double, i64, C.c_double, i64,
                               f64p, i64, i64, f64p, C.POINTER(i64), vp, i64, vp]),
    'ocb_compress_gram_ws_bytes': (i64, [i64, i64]),
    'ocb_compress_from_gram': (C.c_int, [f64p, i64, i64, C.c_double, i64, C.c_double, i64,
                                         f64p, i64, i64, f64p, C.POINTER(i64), vp, i64, vp]),
    'ocb_p2p_put2d': (C.c_int, [f64p, i64, i64, i64, f64p, i64, vp]),
    'ocb_p2p_sum_peers': (C.c_int, [C.POINTER(vp), i64, i64, f64p, vp]),
    'ocb_adi_set_norm_hook': (C.c_int, [vp, vp]),
    'ocb_adi_ws_bytes': (i64, [i64, i64, i64, i64, C.POINTER(vp)]),
    'ocb_adi_run': (C.c_int, [C.POINTER(vp), C.POINTER(C.c_double), i64, i64, i64,
                              i32p, i32p, f64p, f64p, i64, i64, f64p, i64, i64,
                              i32p, i32p, f64p, i64, C.c_double, f64p, i64, i64,
                              C.POINTER(C.c_double), C.POINTER(i64), vp, i64, vp]),
    'ocb_smw_solve_ws_bytes': (i64, [vp, i64, i64]),
    'ocb_smw_solve': (C.c_int, [vp, i64, f64p, i64, i64, i64, f64p, i64, i64,
                                i32p, i32p, f64p, f64p, i64, i64, vp, i64, vp]),
    'ocb_feedback_ws_bytes': (i64, [i64, i64, i64]),
    'ocb_feedback': (C.c_int, [i32p, i32p, f64p, i64, f64p, i64, i64, f64p, i64, i64,
                               f64p, i64, C.c_double, vp, i64, vp]),
    'ocb_sqnorm': (C.c_int, [f64p, i64, i64, i64, f64p, vp]),
}

_lib = None


class OcbError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('optconpy_b200: CUDA library not built: ' + LIB_PATH +
                          ' (run `python -c "import __graft_entry__ as g; g.build()"`'
                          ' or `make -C optconpy_b200/csrc`)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().ocb_last_error()
        raise OcbError('{0} failed with status {1}: {2}'.format(
            what, rc, msg.decode() if msg else ''))
