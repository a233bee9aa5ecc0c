"""Device-side engine: thin Python over the C ABI (include/b200fit.h).

torch is used ONLY for device memory, streams and (in dist.py) the NCCL process group -- every computation on
the fit path is a kernel of libb200fit.so.  There is no CPU fallback: constructing an Engine without a CUDA
device, or without the built library, raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._abi import B200FitError


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Engine(object):
    """One per GPU.  Replaces the per-process pyspeckit fitter registry + fiteach worker pool
    (multi_v_fit.py:96-123, :363-393)."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise B200FitError("no CUDA device visible: mufasa_b200 has no CPU fallback")
        self.lib = _abi.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        ctx = C.c_void_p()
        rc = self.lib.b200fit_create(C.byref(ctx), self.device.index)
        if rc != 0:
            raise B200FitError("b200fit_create failed (%d)" % rc)
        self.ctx = ctx
        self.num_sms = self.lib.b200fit_num_sms(self.ctx)
        self._ws = torch.zeros(256, dtype=torch.uint8, device=self.device)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.b200fit_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.b200fit_last_error(self.ctx)
            raise B200FitError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def nchan_stride(nchan):
        return (int(nchan) + 31) // 32 * 32

    # -- kernels -------------------------------------------------------------------------------------------
    def gather(self, cube_dev, pix_index_dev=None, flip=False, out=None):
        """[nchan, nplane] fp32 (FITS order, flattened plane) -> [npix, stride] pixel-major rows."""
        assert cube_dev.dtype == torch.float32 and cube_dev.is_contiguous() and cube_dev.dim() == 2
        nchan, nplane = cube_dev.shape
        npix = nplane if pix_index_dev is None else int(pix_index_dev.numel())
        stride = self.nchan_stride(nchan)
        if out is None:
            out = torch.empty((npix, stride), dtype=torch.float32, device=self.device)
        if pix_index_dev is not None:
            assert pix_index_dev.dtype == torch.int64 and pix_index_dev.is_contiguous()
        rc = self.lib.b200fit_gather_spectra(self.ctx, _ptr(cube_dev), nchan, nplane, _ptr(pix_index_dev), npix,
                                             1 if flip else 0, _ptr(out), stride, self._stream())
        self._check(rc, "b200fit_gather_spectra")
        return out

    def fit(self, prob, spectra, guesses, err=None, want_counts=True):
        """Batched LM fit.  spectra [npix, stride] fp32, guesses [npix, P] fp64 (device).  Returns a dict of
        device tensors: params, perr [npix, P]; chi2, err_used [npix]; status [npix] int32; counts [npix, 4]."""
        npix, P = int(prob.npix), 4 * int(prob.ncomp)
        assert spectra.dtype == torch.float32 and spectra.is_contiguous() and spectra.shape[0] == npix
        assert guesses.dtype == torch.float64 and guesses.is_contiguous() and tuple(guesses.shape) == (npix, P)
        dev = self.device
        out = dict(params=torch.empty((npix, P), dtype=torch.float64, device=dev),
                   perr=torch.empty((npix, P), dtype=torch.float64, device=dev),
                   chi2=torch.empty(npix, dtype=torch.float64, device=dev),
                   err_used=torch.empty(npix, dtype=torch.float64, device=dev),
                   status=torch.empty(npix, dtype=torch.int32, device=dev),
                   counts=torch.empty((npix, 4), dtype=torch.int32, device=dev) if want_counts else None)
        rc = self.lib.b200fit_fit(self.ctx, C.byref(prob), _ptr(spectra), spectra.shape[1], _ptr(guesses), _ptr(err),
                                  _ptr(out["params"]), _ptr(out["perr"]), _ptr(out["chi2"]), _ptr(out["err_used"]),
                                  _ptr(out["status"]), _ptr(out["counts"]), _ptr(self._ws), self._stream())
        self._check(rc, "b200fit_fit")
        return out

    def model(self, prob, params, stride=None):
        """fp64 model rows [npix, stride] (NaN rows where params are all-zero / non-finite)."""
        npix = int(prob.npix)
        stride = self.nchan_stride(prob.nchan) if stride is None else stride
        out = torch.empty((npix, stride), dtype=torch.float64, device=self.device)
        rc = self.lib.b200fit_model(self.ctx, C.byref(prob), _ptr(params), _ptr(out), stride, self._stream())
        self._check(rc, "b200fit_model")
        return out

    def mask_counts(self, prob, params1, params2, model_f32=True):
        out = torch.empty(int(prob.npix), dtype=torch.int32, device=self.device)
        rc = self.lib.b200fit_mask_counts(self.ctx, C.byref(prob), _ptr(params1), _ptr(params2),
                                          1 if model_f32 else 0, _ptr(out), self._stream())
        self._check(rc, "b200fit_mask_counts")
        return out

    def mask_of_pixel(self, prob, params1, params2, pix, model_f32=True):
        out = np.zeros(int(prob.nchan), dtype=np.uint8)
        rc = self.lib.b200fit_mask_of_pixel(self.ctx, C.byref(prob), _ptr(params1), _ptr(params2),
                                            1 if model_f32 else 0, int(pix), out.ctypes.data_as(C.c_void_p),
                                            self._stream())
        self._check(rc, "b200fit_mask_of_pixel")
        return out.astype(bool)

    def aicc(self, prob, spectra, params1, params2, fill_mask=None, expand=20, model_f32=True, thr=(5.0, 5.0, 5.0)):
        """lnk10/lnk20/lnk21, rss, N and the integer ncomp map for every pixel."""
        npix = int(prob.npix)
        dev = self.device
        out = dict(rss=torch.empty((npix, 3), dtype=torch.float64, device=dev),
                   nsamp=torch.empty(npix, dtype=torch.float64, device=dev),
                   lnk=torch.empty((npix, 3), dtype=torch.float64, device=dev),
                   ncomp=torch.empty(npix, dtype=torch.int8, device=dev))
        fm = None
        if fill_mask is not None:
            fm = np.ascontiguousarray(fill_mask, dtype=np.uint8)
        thr_arr = (C.c_double * 3)(*[float(t) for t in thr])
        rc = self.lib.b200fit_aicc(self.ctx, C.byref(prob), _ptr(spectra), spectra.shape[1], _ptr(params1),
                                   _ptr(params2), fm.ctypes.data_as(C.c_void_p) if fm is not None else None,
                                   int(expand), 1 if model_f32 else 0, C.cast(thr_arr, C.c_void_p), _ptr(out["rss"]),
                                   _ptr(out["nsamp"]), _ptr(out["lnk"]), _ptr(out["ncomp"]), self._stream())
        self._check(rc, "b200fit_aicc")
        return out


_engines = {}


def get_engine(device=None):
    """Process-wide engine per device (created lazily)."""
    if not torch.cuda.is_available():
        raise B200FitError("no CUDA device visible: mufasa_b200 has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else int(device)
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]
