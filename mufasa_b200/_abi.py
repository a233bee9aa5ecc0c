"""ctypes binding of libb200fit.so (include/b200fit.h).  No CPU fallback: if the CUDA library is missing or no
GPU is visible, every entry point raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200FIT_LIB", os.path.join(HERE, "csrc", "libb200fit.so"))   # override: A/B experiments

MODEL_IDS = {"nh3_multi_v": 0, "n2hp_multi_v": 1}
MAX_NCHAN = 2048
ABI_VERSION = 5


class Problem(C.Structure):
    """struct b200fit_problem"""
    _fields_ = [("model_id", C.c_int32), ("ncomp", C.c_int32), ("nchan", C.c_int32), ("max_iter", C.c_int32),
                ("npix", C.c_int64), ("v0", C.c_double), ("dv", C.c_double), ("nu_ref_ghz", C.c_double),
                ("lo", C.c_double * 8), ("hi", C.c_double * 8), ("ftol", C.c_double),
                ("signal_cut", C.c_double), ("weight_mode", C.c_int32), ("reserved", C.c_int32)]


def make_problem(fittype, ncomp, nchan, npix, v0, dv, lo, hi, nu_ref_ghz=0.0, max_iter=200, ftol=1e-10,
                 signal_cut=0.0, weight_mode=0):
    from .exceptions import FitTypeError
    if fittype not in MODEL_IDS:
        raise FitTypeError("'{}' is an invalid fittype".format(fittype))
    p = Problem()
    p.model_id = MODEL_IDS[fittype]
    p.ncomp, p.nchan, p.max_iter, p.npix = int(ncomp), int(nchan), int(max_iter), int(npix)
    p.v0, p.dv, p.nu_ref_ghz = float(v0), float(dv), float(nu_ref_ghz)
    for j in range(8):
        p.lo[j] = float(lo[j]) if j < len(lo) else 0.0
        p.hi[j] = float(hi[j]) if j < len(hi) else 0.0
    p.ftol, p.signal_cut, p.weight_mode = float(ftol), float(signal_cut), int(weight_mode)
    return p


class FitArgs(C.Structure):
    """struct b200fit_fit_args"""
    _fields_ = [("prob", C.POINTER(Problem)), ("d_spectra", C.c_void_p), ("nchan_stride", C.c_int64),
                ("d_row_index", C.c_void_p), ("d_guesses", C.c_void_p), ("d_err", C.c_void_p),
                ("d_params", C.c_void_p), ("d_perr", C.c_void_p), ("d_chi2", C.c_void_p), ("d_err_used", C.c_void_p),
                ("d_status", C.c_void_p), ("d_counts", C.c_void_p), ("d_workspace", C.c_void_p),
                ("wait_event", C.c_void_p)]


MAX_BATCH = 8
PACK_FIELDS = 44


class B200FitError(RuntimeError):
    pass


_lib = None

_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_PP = C.POINTER(Problem)
SIGNATURES = {
    "b200fit_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "b200fit_destroy": (None, [_vp]),
    "b200fit_last_error": (C.c_char_p, [_vp]),
    "b200fit_abi_version": (C.c_int, []),
    "b200fit_num_sms": (C.c_int, [_vp]),
    "b200fit_launch_count": (C.c_uint64, [_vp]),
    "b200fit_host_launch_count": (C.c_uint64, [_vp]),
    "b200fit_measure_xu_peak": (C.c_int, [_vp, C.POINTER(C.c_double), _vp]),
    "b200fit_set_profiling": (C.c_int, [_vp, _i32]),
    "b200fit_get_profile": (C.c_int, [_vp, _vp]),
    "b200fit_nchan_stride": (_i64, [_i32]),
    "b200fit_workspace_bytes": (_sz, [_PP]),
    "b200fit_gather_spectra": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "b200fit_fit": (C.c_int, [_vp, _PP, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200fit_fit_batch": (C.c_int, [_vp, C.POINTER(FitArgs), _i32, _vp]),
    "b200fit_model": (C.c_int, [_vp, _PP, _vp, _vp, _i64, _vp]),
    "b200fit_aicc": (C.c_int, [_vp, _PP, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp]),
    "b200fit_aicc_pair": (C.c_int, [_vp, _PP, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp]),
    "b200fit_aicc_pair_masked": (C.c_int, [_vp, _PP, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp,
                                           _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200fit_mask_bits_pair": (C.c_int, [_vp, _PP, _i32, _i32, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "b200fit_mask_argmax": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "b200fit_mask_bits": (C.c_int, [_vp, _PP, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "b200fit_prepare_guesses": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, C.c_double, _vp, _vp]),
    "b200fit_pack_selection": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b200fit_rms_snr": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, C.c_double, _i32, _vp, _vp, _vp, _vp]),
    "b200fit_signal_bits": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, C.c_double, _vp, _vp]),
    "b200fit_bits_morph": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp]),
    "b200fit_bits_window": (C.c_int, [_vp, _i32, _i64, _i64, C.c_double, C.c_double, _vp, _vp, C.c_double, C.c_double,
                                      _vp, _vp]),
    "b200fit_v_at_peak": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp]),
    "b200fit_masked_moments": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp,
                                         _vp]),
    "b200fit_resid_stats": (C.c_int, [_vp, _PP, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b200fit_upload_slab": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp]),
    "b200fit_convolve_planes": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, C.c_double, _vp, _vp]),
    "b200fit_resample_bilinear": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, C.c_double, C.c_double,
                                            C.c_double, C.c_double, _vp]),
}


def load():
    """dlopen libb200fit.so and declare the prototypes.  Raises B200FitError if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200FitError("libb200fit.so is not built (%s); run __graft_entry__.build() -- there is no CPU fallback"
                           % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    experiment = "B200FIT_LIB" in os.environ          # A/B runs against an older build: tolerate missing hooks
    for name, (res, args) in SIGNATURES.items():
        if experiment and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b200fit_abi_version() != ABI_VERSION and not experiment:
        raise B200FitError("ABI version mismatch")
    _lib = lib
    return lib
