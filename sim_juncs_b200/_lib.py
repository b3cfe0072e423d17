"""ctypes binding of libsimjuncs_b200.so (include/sim_juncs_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
simulation is created, the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SJ_LIB=NAME loads lib/libsimjuncs_b200_NAME.so instead: a build variant for A/B measurements (scripts/build_variant.sh)
LIB_PATH = os.path.join(_HERE, "lib", "libsimjuncs_b200%s.so" % ("_" + os.environ["SJ_LIB"] if os.environ.get("SJ_LIB") else ""))

SJ_MAX_POLES = 4
SJ_F64, SJ_F32 = 0, 1
EX, EY, EZ, HX, HY, HZ = range(6)


class SjGrid(C.Structure):
    _fields_ = [("n", C.c_int32 * 3), ("a", C.c_double), ("courant", C.c_double), ("pml_thickness", C.c_double),
                ("pml_R", C.c_double), ("precision", C.c_int32), ("n_sets", C.c_int32), ("kz0", C.c_int32),
                ("kz1", C.c_int32), ("device", C.c_int32)]


class SjPole(C.Structure):
    _fields_ = [("omega0", C.c_double), ("gamma", C.c_double), ("sigma", C.c_double), ("drude", C.c_int32),
                ("pad", C.c_int32)]


class SjMaterial(C.Structure):
    _fields_ = [("eps_inf", C.c_double), ("n_poles", C.c_int32), ("pad", C.c_int32), ("poles", SjPole * SJ_MAX_POLES)]


class SjCsgNode(C.Structure):
    _fields_ = [("type", C.c_int32), ("invert", C.c_int32), ("cmb", C.c_int32), ("child0", C.c_int32),
                ("child1", C.c_int32), ("pad", C.c_int32), ("M", C.c_double * 9), ("p", C.c_double * 7)]


class SjRegion(C.Structure):
    _fields_ = [("root", C.c_int32), ("n_poles", C.c_int32), ("eps", C.c_double), ("poles", SjPole * SJ_MAX_POLES)]


# every symbol include/sim_juncs_b200.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p
DIPOLE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.POINTER(C.c_double))
SYMBOLS = {
    "sj_create": (C.c_int, [C.POINTER(SjGrid), C.POINTER(_vp)]),
    "sj_destroy": (None, [_vp]),
    "sj_last_error": (C.c_char_p, [_vp]),
    "sj_version": (C.c_int, []),
    "sj_set_materials": (C.c_int, [_vp, C.c_int32, C.POINTER(SjMaterial), _u8p, _u8p, _u8p]),
    "sj_rasterize": (C.c_int, [_vp, C.c_double, C.c_int32, C.POINTER(SjCsgNode), C.c_int32, C.POINTER(SjRegion)]),
    "sj_rasterize_smooth": (C.c_int, [_vp, C.c_double, C.c_int32, C.POINTER(SjCsgNode), C.c_int32, C.POINTER(SjRegion),
                                      C.c_int32, C.c_double]),
    "sj_get_material_ids": (C.c_int, [_vp, C.c_int, _u8p]),
    "sj_get_region_masks": (C.c_int, [_vp, C.c_int, _u8p]),
    "sj_get_material_table": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(SjMaterial), C.c_int32]),
    "sj_add_gaussian_source": (C.c_int, [_vp, C.c_int, _dp, _dp] + [C.c_double] * 7 + [C.c_int, _dp]),
    "sj_add_cw_source": (C.c_int, [_vp, C.c_int, _dp, _dp] + [C.c_double] * 7 + [C.c_int, _dp]),
    "sj_add_custom_source": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_double, C.c_double, DIPOLE_FN, _vp, C.c_double, C.c_int, _dp]),
    "sj_sample_at": (C.c_int, [_vp, C.c_int, C.c_int32, _dp, _dp]),
    "sj_last_source_time": (C.c_double, [_vp]),
    "sj_add_monitors": (C.c_int, [_vp, C.c_int, C.c_int32, _dp]),
    "sj_run": (C.c_int, [_vp, C.c_int64, C.c_int32]),
    "sj_sync": (C.c_int, [_vp]),
    "sj_steps_done": (C.c_int64, [_vp]),
    "sj_n_samples": (C.c_int32, [_vp]),
    "sj_dt": (C.c_double, [_vp]),
    "sj_read_monitors": (C.c_int, [_vp, _dp]),
    "sj_read_spectra": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _dp]),
    "sj_extract_cep": (C.c_int, [_vp, C.c_double, _dp]),
    "sj_pass": (C.c_int, [_vp, C.c_int, C.c_int32, C.c_int32, _vp]),
    "sj_tick": (C.c_int, [_vp, _vp]),
    "sj_sample": (C.c_int, [_vp, _vp]),
    "sj_plane_ptr": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int32, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "sj_halo_exchange": (C.c_int, [_vp, _vp, C.c_int]),
    "sj_export_peer": (C.c_int, [_vp, C.c_char_p]),
    "sj_connect_peers": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "sj_connect_local": (C.c_int, [_vp, _vp]),
    "sj_run_group": (C.c_int, [C.POINTER(_vp), C.c_int32, C.c_int64, C.c_int32]),
    "sj_get_field": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "sj_run_timed": (C.c_int, [_vp, C.c_int64, C.c_int32, _dp]),
    "sj_profile_kernels": (C.c_int, [_vp, C.c_int32, _dp]),
    "sj_get_counts": (C.c_int, [_vp, _dp]),
    "sj_plane_costs": (C.c_int, [_vp, _dp]),
    "sj_memory": (C.c_int, [_vp, _dp]),
    "sj_get_stats": (C.c_int, [_vp, C.POINTER(C.c_int64), _dp]),
    "sj_bytes_per_step": (C.c_double, [_vp]),
}

_lib = None


def load():
    """Load the C-ABI library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libsimjuncs_b200.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C sim_juncs_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
