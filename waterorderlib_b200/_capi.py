"""ctypes binding of include/wol_capi.h.  There is no CPU fallback: if libwol.so is missing or a call
fails, an exception is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwol.so")

WOL_OK = 0
WOL_F64, WOL_F32 = 0, 1
WOL_PREC_FP64, WOL_PREC_FP32 = 0, 1
WOL_NSTATS = 8
WOL_TABLE_EXTRA = 8
STAT_NAMES = ("q_sum", "q_sumsq", "n_centres", "tet_count", "tet_cos", "tet_cossq", "n_angles", "n_neigh")

c_i32 = ctypes.c_int32
c_vp = ctypes.c_void_p


class Q3bArgs(ctypes.Structure):
    """struct wol_q3b_args (include/wol_capi.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("precision", c_i32), ("n_frames", c_i32), ("n_pos", c_i32),
        ("n_centres", c_i32), ("centre_dtype", c_i32), ("centres", c_vp), ("box", c_vp), ("workspace", c_vp),
        ("workspace_bytes", ctypes.c_size_t), ("nc", c_i32 * 3), ("hist_per_frame", c_i32),
        ("edge_min", ctypes.c_double), ("box_max", ctypes.c_double), ("low3", ctypes.c_double), ("high3", ctypes.c_double),
        ("lowq", ctypes.c_double), ("highq", ctypes.c_double), ("do_q", c_i32), ("do_3body", c_i32),
        ("nbins", c_i32), ("q_nbins", c_i32), ("hist_lo", ctypes.c_double), ("hist_hi", ctypes.c_double),
        ("angle_table", c_vp), ("q", c_vp), ("nn_idx", c_vp), ("n3", c_vp), ("ang_hist", c_vp), ("q_hist", c_vp),
        ("frame_stats", c_vp), ("timing_event_begin", c_vp), ("timing_event_end", c_vp), ("n_valid", c_vp),
    ]


class HbondArgs(ctypes.Structure):
    """struct wol_hbond_args (include/wol_capi.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("n_frames", c_i32), ("n_acc", c_i32), ("n_don", c_i32), ("acc_dtype", c_i32),
        ("donh_dtype", c_i32), ("acc", c_vp), ("donh", c_vp), ("box", c_vp), ("workspace", c_vp),
        ("workspace_bytes", ctypes.c_size_t), ("nc", c_i32 * 3), ("pair_capacity", ctypes.c_uint32),
        ("edge_min", ctypes.c_double), ("dist_cut", ctypes.c_double), ("ang_cut", ctypes.c_double),
        ("acc_count", c_vp), ("don_count", c_vp), ("dense", c_vp), ("pairs", c_vp), ("pair_counter", c_vp),
    ]


c_f64 = ctypes.c_double
c_i64 = ctypes.c_int64
_NC = ctypes.POINTER(c_i32 * 3)

# name -> (restype, argtypes); every symbol include/wol_capi.h declares
SIGNATURES = {
    "wol_version": (ctypes.c_char_p, []),
    "wol_last_error": (ctypes.c_char_p, []),
    "wol_abi_version": (ctypes.c_int, []),
    "wol_last_launch_count": (ctypes.c_int, []),
    "wol_plan_grid": (ctypes.c_int, [c_vp, c_i32, ctypes.c_double, ctypes.POINTER(c_i32 * 3), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_double)]),
    "wol_workspace_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32, ctypes.POINTER(c_i32 * 3)]),
    "wol_effective_box": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, ctypes.c_double, c_vp, c_vp, c_vp]),
    "wol_cell_build": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, ctypes.POINTER(c_i32 * 3), c_i32, c_vp, ctypes.c_size_t, c_vp]),
    "wol_angle_table": (ctypes.c_int, [ctypes.c_double, ctypes.c_double, c_i32, ctypes.c_double, ctypes.c_double, c_vp]),
    "wol_q3b_frames": (ctypes.c_int, [ctypes.POINTER(Q3bArgs), c_vp]),
    "wol_angle_offsets": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "wol_angles_fill": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, _NC, c_f64, c_f64, c_f64, c_vp,
                                       ctypes.c_size_t, c_vp, c_vp, c_vp]),
    "wol_histogram": (ctypes.c_int, [c_vp, c_i64, c_f64, c_f64, c_i32, c_vp, c_f64, c_f64, c_vp, c_vp]),
    "wol_neighbor_matrix": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_f64, c_f64, c_vp, c_vp]),
    "wol_reimage": (ctypes.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp]),
    "wol_tetracosang": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "wol_lsi": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, _NC, c_f64, c_f64, c_f64, c_vp, ctypes.c_size_t, c_vp, c_vp, c_vp]),
    "wol_hbond_counts": (ctypes.c_int, [ctypes.POINTER(HbondArgs), c_vp]),
    "wol_hbond_locations": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "wol_shell_mask": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, _NC, c_f64, c_f64, c_f64, c_vp,
                                      ctypes.c_size_t, c_vp, c_vp]),
    "wol_willard_density": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, _NC, c_f64, c_f64, c_vp,
                                           ctypes.c_size_t, c_vp, c_vp, c_vp]),
    "wol_density_field": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f64, c_vp, c_i32, _NC, c_f64, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "wol_neighbors_csr": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, _NC, c_f64, c_f64, c_f64, c_vp, ctypes.c_size_t, c_vp, c_vp, c_vp,
                                         c_i64, c_vp]),
    "wol_water_orient": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "wol_bin_on_grid": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f64, c_vp, c_vp]),
    "wol_hist_allreduce": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_i32, c_vp]),
    "wol_fma_probe": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_vp]),
    "wol_iso_face_scratch_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32]),
    "wol_iso_faces": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_f64, c_vp, c_vp, ctypes.c_size_t, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "wol_radial_dist_plane": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_f64, c_i32, c_f64, c_vp, c_vp, c_vp]),
    "wol_histogram2d": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp]),
    "wol_iso_scratch_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32]),
    "wol_iso_points": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f64, c_vp, ctypes.c_size_t, c_vp, ctypes.c_int64, c_vp, c_vp]),
    "wol_interface_water": (ctypes.c_int, [c_vp, c_i32, c_vp, c_vp, c_i32, c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "wol_profile_bins": (ctypes.c_int, [c_vp, c_vp, c_i64, c_f64, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "wol_angle_table_ceil": (ctypes.c_int, [c_f64, c_i32, c_vp]),
    "wol_histrr3b": (ctypes.c_int, [c_vp, c_i32, _NC, c_f64, c_f64, c_i32, c_f64, c_i32, c_vp, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "wol_pair_hist": (ctypes.c_int, [c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, _NC, c_f64, c_f64, c_i32, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "wol_components": (ctypes.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "wol_psi": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, _NC, c_f64, c_f64, c_f64, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "wol_status": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, ctypes.POINTER(c_i32 * 3), c_vp, ctypes.POINTER(c_i32 * 4)]),
}

_LIB = None


class WolError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises if it has not been built -- the product has no other path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise WolError("libwol.so is not built (run `python -m waterorderlib_b200.build`); there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc, what):
    if rc != WOL_OK:
        msg = lib().wol_last_error().decode("utf-8", "replace")
        exc = ValueError if rc in (-1, -2) else WolError
        raise exc("%s failed (code %d): %s" % (what, rc, msg))
