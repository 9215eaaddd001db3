"""ctypes binding of libgofpid_b200.so (the C ABI declared in include/gofpid_b200.h).

The library is the product: there is no CPU fallback.  If the shared object is missing the
import of this module fails loudly, and every compute entry point returns GOFPID_ERR_CUDA when
no CUDA device is present.
"""
import ctypes as ct
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# (GOFPID_LIB: file name of a variant build next to the default library -- kernel experiments only)
LIB_PATH = os.path.join(_HERE, os.environ.get("GOFPID_LIB") or "libgofpid_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM, ERR_OVERFLOW = -1, -2, -3, -4, -5
BLUR_NONE, BLUR_GAUSSIAN, BLUR_BOX = 0, 1, 2
DETECT_FD, DETECT_MOG2, DETECT_KNN = 0, 1, 2
MORPH_ERODE, MORPH_DILATE = 0, 1
MAX_MORPH_OPS = 8
MAX_KSIZE = 31
MAX_SE = 31


class MorphOp(ct.Structure):
    _fields_ = [("op", ct.c_int32), ("kh", ct.c_int32), ("kw", ct.c_int32),
                ("kernel", ct.POINTER(ct.c_uint8))]


class Config(ct.Structure):
    _fields_ = [
        ("blur_kind", ct.c_int32), ("blur_kw", ct.c_int32), ("blur_kh", ct.c_int32),
        ("blur_sigma", ct.c_double),
        ("detector", ct.c_int32),
        ("fd_threshold", ct.c_double),
        ("mog2_history", ct.c_int32),
        ("mog2_var_threshold", ct.c_float), ("mog2_var_threshold_gen", ct.c_float),
        ("mog2_background_ratio", ct.c_float),
        ("mog2_var_init", ct.c_float), ("mog2_var_min", ct.c_float), ("mog2_var_max", ct.c_float),
        ("mog2_ct", ct.c_float),
        ("knn_history", ct.c_int32), ("knn_k", ct.c_int32),
        ("knn_dist2_threshold", ct.c_float),
        ("knn_rng_seed", ct.c_uint64),
        ("detect_shadows", ct.c_int32), ("shadow_value", ct.c_int32), ("shadow_tau", ct.c_float),
        ("n_morph", ct.c_int32),
        ("morph", MorphOp * MAX_MORPH_OPS),
        ("find_blobs", ct.c_int32),
        ("blob_area_min", ct.c_double),
        ("max_blobs", ct.c_int32), ("max_kept", ct.c_int32), ("max_points", ct.c_int32),
        ("convert", ct.c_int32),
        ("depth", ct.c_int32), ("mog2_group", ct.c_int32),
    ]


class Blob(ct.Structure):
    _fields_ = [("first", ct.c_int32), ("x0", ct.c_int32), ("y0", ct.c_int32), ("x1", ct.c_int32),
                ("y1", ct.c_int32), ("area_px", ct.c_int32), ("n_points", ct.c_int32),
                ("point_offset", ct.c_int32), ("nested", ct.c_int32), ("reserved", ct.c_int32),
                ("a00", ct.c_int64), ("a10", ct.c_int64),
                ("a01", ct.c_int64)]


class FrameInfo(ct.Structure):
    _fields_ = [("n_components", ct.c_int32), ("n_candidates", ct.c_int32), ("n_kept", ct.c_int32),
                ("n_points", ct.c_int32), ("overflow", ct.c_int32), ("frame_index", ct.c_int32),
                ("reserved", ct.c_int32 * 2)]


class TrackerConfig(ct.Structure):
    _fields_ = [("H", ct.c_int32), ("persp_height", ct.POINTER(ct.c_double)), ("persp_area", ct.POINTER(ct.c_double)),
                ("perimeter", ct.POINTER(ct.c_int32)), ("n_perimeter", ct.c_int32),
                ("tracking_factor", ct.c_double), ("absence_max", ct.c_int32), ("anchor_center", ct.c_int32),
                ("perspective_coeff", ct.c_double), ("presence_min", ct.c_double), ("distance_min", ct.c_double),
                ("keep_alarm", ct.c_int32), ("duration_min", ct.c_double)]


class Track(ct.Structure):
    _fields_ = [("anchor", ct.c_int16 * 2), ("bottom", ct.c_int16 * 2), ("center", ct.c_int16 * 2),
                ("center_first", ct.c_int16 * 2), ("presence", ct.c_int32), ("absence", ct.c_int32),
                ("filtered", ct.c_int32), ("already_in_alarm", ct.c_int32),
                ("x0", ct.c_int32), ("y0", ct.c_int32), ("x1", ct.c_int32), ("y1", ct.c_int32), ("n_points", ct.c_int32)]


# name -> (restype, argtypes); must list every GOFPID_API symbol of include/gofpid_b200.h
_u8p = ct.POINTER(ct.c_uint8)
_i32p = ct.POINTER(ct.c_int32)
_f32p = ct.POINTER(ct.c_float)
_vp = ct.c_void_p
_sz = ct.c_size_t
SIGNATURES = {
    "gofpid_version": (ct.c_char_p, []),
    "gofpid_last_error": (ct.c_char_p, []),
    "gofpid_device_count": (ct.c_int, []),
    "gofpid_config_default": (None, [ct.POINTER(Config)]),
    "gofpid_host_alloc": (ct.c_int, [_sz, ct.c_int, ct.POINTER(_vp)]),
    "gofpid_host_copy2d": (ct.c_int, [_vp, _sz, _vp, _sz, _sz, _sz, ct.c_int]),
    "gofpid_host_free": (ct.c_int, [_vp]),
    "gofpid_gaussian_taps_q8": (ct.c_int, [ct.c_int, ct.c_double, ct.POINTER(ct.c_uint16)]),
    "gofpid_pipeline_create": (ct.c_int, [ct.POINTER(Config), ct.c_int, ct.c_int, ct.c_int, ct.c_int,
                                          ct.c_int, ct.POINTER(_vp)]),
    "gofpid_pipeline_destroy": (ct.c_int, [_vp]),
    "gofpid_pipeline_reset": (ct.c_int, [_vp]),
    "gofpid_pipeline_process_device": (ct.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "gofpid_pipeline_process_host": (ct.c_int, [_vp, _vp, _sz, _sz]),
    "gofpid_pipeline_submit_host": (ct.c_int, [_vp, _vp, _sz, _sz]),
    "gofpid_pipeline_process_host_sequence": (ct.c_int, [_vp, _vp, ct.c_int, _sz, _sz, _sz, ct.POINTER(Blob), ct.c_int,
                                                         ct.POINTER(ct.c_int32), ct.c_int, ct.POINTER(ct.c_int32)]),
    "gofpid_pipeline_process_device_sequence": (ct.c_int, [_vp, _vp, ct.c_int, _sz, _sz, _sz, _vp, ct.POINTER(Blob),
                                                           ct.c_int, ct.POINTER(ct.c_int32), ct.c_int,
                                                           ct.POINTER(ct.c_int32)]),
    "gofpid_pipeline_sequence_results": (ct.c_int, [_vp, ct.POINTER(Blob), ct.c_int, ct.POINTER(ct.c_int32), ct.c_int]),
    "gofpid_pipeline_join": (ct.c_int, [_vp, _vp]),
    "gofpid_pipeline_group": (ct.c_int, [_vp]),
    "gofpid_pipeline_select": (ct.c_int, [_vp, ct.c_int]),
    "gofpid_pipeline_depth": (ct.c_int, [_vp]),
    "gofpid_pipeline_frame_info": (ct.c_int, [_vp, ct.c_int, ct.POINTER(FrameInfo)]),
    "gofpid_pipeline_blobs": (ct.c_int, [_vp, ct.c_int, ct.POINTER(Blob), ct.c_int, _i32p, ct.c_int,
                                         ct.POINTER(ct.c_int), ct.POINTER(ct.c_int)]),
    "gofpid_pipeline_mask": (ct.c_int, [_vp, ct.c_int, _vp, _sz, ct.c_int]),
    "gofpid_pipeline_labels": (ct.c_int, [_vp, ct.c_int, _vp, _sz, ct.c_int]),
    "gofpid_pipeline_components": (ct.c_int, [_vp, ct.c_int, _i32p, ct.c_int, ct.POINTER(ct.c_int)]),
    "gofpid_pipeline_mog2_model": (ct.c_int, [_vp, ct.c_int, _f32p, _f32p, _f32p, _u8p]),
    "gofpid_pipeline_knn_model": (ct.c_int, [_vp, ct.c_int, _u8p, _u8p, _u8p, _i32p,
                                             ct.POINTER(ct.c_uint64)]),
    "gofpid_pipeline_fd_model": (ct.c_int, [_vp, ct.c_int, _u8p, ct.POINTER(ct.c_int)]),
    "gofpid_pipeline_mog2_mean_modes": (ct.c_int, [_vp, ct.POINTER(ct.c_double)]),
    "gofpid_pipeline_set_profiling": (ct.c_int, [_vp, ct.c_int]),
    "gofpid_pipeline_stage_ms": (ct.c_int, [_vp, _f32p]),
    "gofpid_pipeline_launch_count": (ct.c_int, [_vp]),
    "gofpid_pipeline_launch_total": (ct.c_int, [_vp, ct.POINTER(ct.c_longlong)]),
    "gofpid_tracker_create": (ct.c_int, [ct.POINTER(TrackerConfig), ct.POINTER(_vp)]),
    "gofpid_tracker_destroy": (ct.c_int, [_vp]),
    "gofpid_tracker_reset": (ct.c_int, [_vp]),
    "gofpid_tracker_step": (ct.c_int, [_vp, _vp, ct.c_int, _vp, ct.POINTER(ct.c_int)]),
    "gofpid_tracker_step_prepared": (ct.c_int, [_vp, ct.c_int, _vp, _vp, _vp, _vp, _vp, _vp, ct.POINTER(ct.c_int)]),
    "gofpid_tracker_run": (ct.c_int, [_vp, _vp, _vp, _vp, ct.c_int, ct.c_int, ct.c_int, _vp]),
    "gofpid_tracker_count": (ct.c_int, [_vp]),
    "gofpid_tracker_get": (ct.c_int, [_vp, ct.c_int, ct.POINTER(Track)]),
    "gofpid_tracker_contour": (ct.c_int, [_vp, ct.c_int, _vp, ct.c_int]),
    "gofpid_tracker_alarm_state": (ct.c_int, [_vp, ct.POINTER(ct.c_int), ct.POINTER(ct.c_double)]),
    "gofpid_blur_u8": (ct.c_int, [_vp, _sz, _vp, _sz, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int,
                                  ct.c_int, ct.c_double, _vp]),
    "gofpid_morph_u8": (ct.c_int, [_vp, _sz, _vp, _sz, ct.c_int, ct.c_int, ct.POINTER(MorphOp), ct.c_int,
                                   _vp]),
    "gofpid_morph_levels_u8": (ct.c_int, [_vp, _sz, _vp, _sz, ct.c_int, ct.c_int, ct.POINTER(MorphOp), ct.c_int,
                                          ct.POINTER(ct.c_int), ct.c_int, _vp]),
    "gofpid_ccl_u8": (ct.c_int, [_vp, _sz, _vp, _sz, ct.c_int, ct.c_int, _vp]),
}


def build(force=False):
    """Compile libgofpid_b200.so for sm_100a with nvcc (in-tree, idempotent)."""
    args = ["make", "-C", CSRC, "-j4"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Raises ImportError (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libgofpid_b200.so is not built (%s missing): run `python -c \"import "
                "__graft_entry__ as g; g.build()\"` or `make -C pygofpid_b200/csrc`. "
                "pygofpid_b200 has no CPU fallback." % LIB_PATH)
        L = ct.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)  # AttributeError if the ABI and the binding diverge
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class GofpidError(RuntimeError):
    """CUDA / capacity failures reported by the library."""


def last_error():
    return lib().gofpid_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a status code to the exception type the reference raises for the same mistake."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ERR_OVERFLOW:
        raise GofpidError("capacity exceeded: " + msg)
    raise GofpidError(msg)
