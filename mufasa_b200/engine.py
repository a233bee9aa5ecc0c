"""Device-side engine: thin Python over the C ABI (include/b200fit.h).

torch is used ONLY for device memory, streams and (in dist.py) the NCCL process group -- every computation on
the fit path is a kernel of libb200fit.so.  There is no CPU fallback: constructing an Engine without a CUDA
device, or without the built library, raises.
"""
import ctypes as C
import os

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
        self._ws = torch.zeros(4096, dtype=torch.uint8, device=self.device)

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

    def workspace(self, prob):
        """Device scratch of b200fit_workspace_bytes(prob) bytes (grown on demand, reused across calls)."""
        need = int(self.lib.b200fit_workspace_bytes(C.byref(prob)))
        if self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def pinned_stage(self, shape, dtype=torch.float64):
        """Pinned host staging buffer for D2H copies, cached per shape (page-locking costs ~10 ms per call)."""
        cache = self.__dict__.setdefault("_pinned", {})
        key = (tuple(shape), dtype)         # a leading string in ``shape`` names a separate buffer of that shape
        if key not in cache:
            if len(cache) > 16:
                cache.clear()
            cache[key] = torch.empty([d for d in shape if not isinstance(d, str)], dtype=dtype).pin_memory()
        return cache[key]

    def side_stream(self, name):
        """A second CUDA stream of this engine (created on first use), e.g. for copies that must not queue behind
        the main stream's work."""
        streams = self.__dict__.setdefault("_side_streams", {})
        if name not in streams:
            streams[name] = torch.cuda.Stream(device=self.device)
        return streams[name]

    @staticmethod
    def nchan_stride(nchan):
        return (int(nchan) + 31) // 32 * 32

    # -- measurement hooks ---------------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.b200fit_launch_count(self.ctx))

    def host_launch_count(self):
        """Launches the host issued (kernels outside graphs + graph launches)."""
        return int(self.lib.b200fit_host_launch_count(self.ctx))

    def measure_xu_peak(self):
        """ex2.approx operations per second of this GPU (b200fit_measure_xu_peak micro-benchmark)."""
        out = C.c_double(0.0)
        self._check(self.lib.b200fit_measure_xu_peak(self.ctx, C.byref(out), self._stream()), "b200fit_measure_xu_peak")
        return float(out.value)

    def set_profiling(self, on):
        self._profiling = bool(on)          # profiled fits run undivided (one event list per fit)
        self._check(self.lib.b200fit_set_profiling(self.ctx, 1 if on else 0), "b200fit_set_profiling")

    def get_profile(self):
        """dict(prologue_ms, eval_ms, solve_ms, rounds) of the last profiled fit."""
        out = (C.c_double * 4)()
        self._check(self.lib.b200fit_get_profile(self.ctx, C.cast(out, C.c_void_p)), "b200fit_get_profile")
        return dict(prologue_ms=out[0], eval_ms=out[1], solve_ms=out[2], rounds=int(out[3]))

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

    def upload_slab(self, host_cube, first, count, dst):
        """Pinned host cube [nchan, nplane] fp32 -> dst [nchan, count]: elements [first, first + count) of every
        plane, one 2-D copy on the current stream (b200fit_upload_slab)."""
        assert host_cube.dtype == torch.float32 and host_cube.is_contiguous() and host_cube.dim() == 2
        assert dst.dtype == torch.float32 and dst.is_contiguous() and tuple(dst.shape) == (host_cube.shape[0], count)
        rc = self.lib.b200fit_upload_slab(self.ctx, C.c_void_p(host_cube.data_ptr()), int(host_cube.shape[0]),
                                          int(host_cube.shape[1]), int(first), int(count), _ptr(dst), self._stream())
        self._check(rc, "b200fit_upload_slab")

    def fit(self, prob, spectra, guesses, err=None, want_counts=True, row_index=None, out=None):
        """Batched LM fit.  spectra [nrows, stride] fp32, guesses [npix, P] fp64 (device); ``row_index`` [npix] int64
        maps pixel k to its spectrum row (None: row k).  Returns a dict of device tensors: params, perr [npix, P];
        chi2, err_used [npix]; status [npix] int32; counts [npix, 4]."""
        npix, P = int(prob.npix), 4 * int(prob.ncomp)
        assert spectra.dtype == torch.float32 and spectra.is_contiguous()
        assert row_index is not None or spectra.shape[0] == npix
        assert guesses.dtype == torch.float64 and guesses.is_contiguous() and tuple(guesses.shape) == (npix, P)
        if row_index is not None:
            assert row_index.dtype == torch.int64 and row_index.is_contiguous() and row_index.numel() == npix
        dev = self.device
        if out is None:
            out = dict(params=torch.empty((npix, P), dtype=torch.float64, device=dev),
                       perr=torch.empty((npix, P), dtype=torch.float64, device=dev),
                       chi2=torch.empty(npix, dtype=torch.float64, device=dev),
                       err_used=torch.empty(npix, dtype=torch.float64, device=dev),
                       status=torch.empty(npix, dtype=torch.int32, device=dev),
                       counts=torch.empty((npix, 4), dtype=torch.int32, device=dev) if want_counts else None)
        if npix >= 2 * self.SPLIT_MIN_PIXELS and not getattr(self, "_profiling", False):
            # a large fit runs as a few independent parts in flight together (see _split_jobs)
            return self.fit_batch([dict(prob=prob, spectra=spectra, guesses=guesses, err=err, row_index=row_index,
                                        out=out)])[0]
        rc = self.lib.b200fit_fit(self.ctx, C.byref(prob), _ptr(spectra), spectra.shape[1], _ptr(row_index),
                                  _ptr(guesses), _ptr(err),
                                  _ptr(out["params"]), _ptr(out["perr"]), _ptr(out["chi2"]), _ptr(out["err_used"]),
                                  _ptr(out["status"]), _ptr(out["counts"]), _ptr(self.workspace(prob)),
                                  self._stream())
        self._check(rc, "b200fit_fit")
        return out

    def alloc_fit_out(self, prob, want_counts=True):
        npix, P = int(prob.npix), 4 * int(prob.ncomp)
        dev = self.device
        return dict(params=torch.empty((npix, P), dtype=torch.float64, device=dev),
                    perr=torch.empty((npix, P), dtype=torch.float64, device=dev),
                    chi2=torch.empty(npix, dtype=torch.float64, device=dev),
                    err_used=torch.empty(npix, dtype=torch.float64, device=dev),
                    status=torch.empty(npix, dtype=torch.int32, device=dev),
                    counts=torch.empty((npix, 4), dtype=torch.int32, device=dev) if want_counts else None)

    def fit_batch(self, jobs):
        """Several fits in flight at once (b200fit_fit_batch): ``jobs`` = list of dicts with keys prob, spectra,
        guesses and optionally err, row_index, out.  Typically the 1- and 2-component fits of the same rows: the
        latency-bound late rounds of one overlap the other's work.  Returns the list of result dicts."""
        if len(jobs) > _abi.MAX_BATCH:
            raise B200FitError("at most %d fits per batch" % _abi.MAX_BATCH)
        if not hasattr(self, "_ws_batch"):
            self._ws_batch = [None] * _abi.MAX_BATCH
        results = []
        for job in jobs:
            prob, spectra, guesses = job["prob"], job["spectra"], job["guesses"]
            npix, P = int(prob.npix), 4 * int(prob.ncomp)
            assert spectra.dtype == torch.float32 and spectra.is_contiguous()
            assert guesses.dtype == torch.float64 and guesses.is_contiguous() and tuple(guesses.shape) == (npix, P)
            assert job.get("row_index") is not None or spectra.shape[0] == npix
            results.append(job.get("out") or self.alloc_fit_out(prob))
        for job in jobs:                      # inputs uploaded on another stream (PCube.fiteach)
            if job.get("ready") is not None:
                torch.cuda.current_stream(self.device).wait_event(job["ready"])
        jobs = self._split_jobs(jobs, results)
        n = len(jobs)
        arr = (_abi.FitArgs * max(n, 1))()
        outs, keep = [], []
        for j, job in enumerate(jobs):
            prob, spectra, guesses = job["prob"], job["spectra"], job["guesses"]
            npix, P = int(prob.npix), 4 * int(prob.ncomp)
            row_index = job.get("row_index")
            out = job["out"]
            need = int(self.lib.b200fit_workspace_bytes(C.byref(prob)))
            ws = self._ws_batch[j]
            if ws is None or ws.numel() < need:
                self._ws_batch[j] = None
                ws = self._ws_batch[j] = torch.empty(need, dtype=torch.uint8, device=self.device)
            a = arr[j]
            a.prob = C.pointer(prob)
            a.d_spectra, a.nchan_stride = spectra.data_ptr(), spectra.shape[1]
            a.d_row_index = row_index.data_ptr() if row_index is not None else None
            a.d_guesses = guesses.data_ptr()
            a.d_err = job["err"].data_ptr() if job.get("err") is not None else None
            a.d_params, a.d_perr = out["params"].data_ptr(), out["perr"].data_ptr()
            a.d_chi2, a.d_err_used = out["chi2"].data_ptr(), out["err_used"].data_ptr()
            a.d_status = out["status"].data_ptr()
            a.d_counts = out["counts"].data_ptr() if out.get("counts") is not None else None
            a.d_workspace = ws.data_ptr()
            a.wait_event = job["wait"].cuda_event if job.get("wait") is not None else None
            outs.append(out)
            keep.append((prob, spectra, guesses, row_index))
        rc = self.lib.b200fit_fit_batch(self.ctx, arr, n, self._stream())
        self._check(rc, "b200fit_fit_batch")
        return results

    SPLIT_MIN_PIXELS = 16384

    def _split_jobs(self, jobs, results):
        """Free slots of the batch are used to run the largest fits as independent, equal parts of their pixels
        (contiguous views of the same inputs and outputs).  Pixels do not interact, so the results are those of the
        undivided fit; two half-sized round loops interleave on the GPU where one full-sized loop leaves it partly
        idle between its dependent kernels (measured: 2-component fit of 65 536 spectra 34.7 -> 32.7 ms)."""
        # slab events matter only while the upload is in flight; a resident cube takes the generic split below
        slabbed = [j for j in jobs if j.get("slab_ready") is not None
                   and not all(ev.query() for ev in j["slab_ready"][1])]
        if slabbed and not getattr(self, "_profiling", False):
            return self._split_by_slab(jobs, results)
        for j in slabbed:                     # profiled fits run undivided: order them after the whole upload
            torch.cuda.current_stream(self.device).wait_event(j["slab_ready"][1][-1])
        limit = min(_abi.MAX_BATCH, int(os.environ.get("B200FIT_SPLIT_LIMIT", 3)))   # measured: 3 beats 2 and 4
        parts = [1] * len(jobs)
        weight = [int(j["prob"].npix) * int(j["prob"].ncomp) ** 2 for j in jobs]
        while sum(parts) < limit and not getattr(self, "_profiling", False):
            ok = [i for i in range(len(jobs)) if int(jobs[i]["prob"].npix) // (parts[i] + 1) >= self.SPLIT_MIN_PIXELS]
            if not ok:
                break
            parts[max(ok, key=lambda i: weight[i] / parts[i])] += 1
        work = []
        for job, out, k in zip(jobs, results, parts):
            prob = job["prob"]
            npix = int(prob.npix)
            if k == 1:
                work.append(dict(job, out=out))
                continue
            ri = job.get("row_index")
            for q in range(k):
                a, b = q * npix // k, (q + 1) * npix // k
                p = _abi.Problem()
                C.memmove(C.byref(p), C.byref(prob), C.sizeof(p))
                p.npix = b - a
                work.append(dict(prob=p, guesses=job["guesses"][a:b],
                                 spectra=job["spectra"] if ri is not None else job["spectra"][a:b],
                                 row_index=None if ri is None else ri[a:b],
                                 err=None if job.get("err") is None else job["err"][a:b],
                                 out={key: (None if t is None else t[a:b]) for key, t in out.items()}))
        return work

    def _split_by_slab(self, jobs, results):
        """Jobs whose spectra are still being uploaded slab by slab (PCube.rows_on_device): a job is cut at the slab
        boundaries when the batch has room, and every part waits only for the event of the last slab it reads --
        the first rows are being fitted while the rest of the cube crosses PCIe."""
        nslab = max(len(j["slab_ready"][0]) for j in jobs if j.get("slab_ready") is not None)
        cut = len(jobs) * nslab <= _abi.MAX_BATCH and all(
            int(j["prob"].npix) >= 2 * self.SPLIT_MIN_PIXELS for j in jobs if j.get("slab_ready") is not None)
        work = []
        for job, out in zip(jobs, results):
            prob, npix = job["prob"], int(job["prob"].npix)
            sr = job.get("slab_ready")
            if sr is None:
                work.append(dict(job, out=out))
                continue
            bounds, events = sr
            pix = job.get("pix")                      # ascending pixel indices of a masked fit, or None / all
            if job.get("row_index") is None:
                ends = [min(b, npix) for b in bounds]
            else:
                ends = [int(np.searchsorted(pix, b)) for b in bounds]
            if not cut:
                last = min(k for k in range(len(ends)) if ends[k] >= npix)   # the slab holding its last pixel
                work.append(dict(job, out=out, wait=events[last]))
                continue
            ri, a = job.get("row_index"), 0
            for k, b in enumerate(ends):
                if b <= a:
                    continue
                p = _abi.Problem()
                C.memmove(C.byref(p), C.byref(prob), C.sizeof(p))
                p.npix = b - a
                work.append(dict(prob=p, guesses=job["guesses"][a:b], wait=events[k],
                                 spectra=job["spectra"] if ri is not None else job["spectra"][a:b],
                                 row_index=None if ri is None else ri[a:b],
                                 err=None if job.get("err") is None else job["err"][a:b],
                                 out={key: (None if t is None else t[a:b]) for key, t in out.items()}))
                a = b
        return work

    def model(self, prob, params, stride=None):
        """fp64 model rows [npix, stride] (NaN rows where params are all-zero / non-finite)."""
        npix = int(prob.npix)
        stride = self.nchan_stride(prob.nchan) if stride is None else stride
        out = torch.empty((npix, stride), dtype=torch.float64, device=self.device)
        rc = self.lib.b200fit_model(self.ctx, C.byref(prob), _ptr(params), _ptr(out), stride, self._stream())
        self._check(rc, "b200fit_model")
        return out

    # -- model selection ---------------------------------------------------------------------------------------
    def aicc_alloc(self, npix):
        dev = self.device
        return dict(rss=torch.empty((npix, 3), dtype=torch.float64, device=dev),
                    nsamp=torch.empty(npix, dtype=torch.float64, device=dev),
                    lnk=torch.empty((npix, 3), dtype=torch.float64, device=dev),
                    ncomp=torch.empty(npix, dtype=torch.int8, device=dev),
                    mask_counts=torch.empty(npix, dtype=torch.int32, device=dev))

    def aicc_pass(self, prob, spectra, params1, params2, out, fill_bits=None, only_empty=False, expand=20,
                  model_f32=True, thr=(5.0, 5.0, 5.0), row_index=None, ncomp_pair=(1, 2), mask_bits=None):
        """``mask_bits`` [npix, 64] int32 device tensor (bit b of word w = channel 32 w + b): the caller's spectral
        mask instead of the union of the two model masks (b200fit_aicc_pair_masked)."""
        thr_arr = (C.c_double * 3)(*[float(t) for t in thr])
        if mask_bits is not None:
            assert mask_bits.dtype == torch.int32 and mask_bits.is_contiguous() and tuple(mask_bits.shape) == (
                int(prob.npix), 64)
        rc = self.lib.b200fit_aicc_pair_masked(self.ctx, C.byref(prob), int(ncomp_pair[0]), int(ncomp_pair[1]),
                                   _ptr(spectra), spectra.shape[1], _ptr(row_index),
                                   _ptr(params1), _ptr(params2), _ptr(mask_bits), _ptr(fill_bits),
                                   1 if only_empty else 0, int(expand),
                                   1 if model_f32 else 0, C.cast(thr_arr, C.c_void_p), _ptr(out["rss"]),
                                   _ptr(out["nsamp"]), _ptr(out["lnk"]), _ptr(out["ncomp"]), _ptr(out["mask_counts"]),
                                   self._stream())
        self._check(rc, "b200fit_aicc")
        return out

    def mask_argmax(self, mask_counts, index_offset=0):
        """device int64[2] = {largest mask count, index_offset + first pixel holding it}."""
        best = torch.empty(2, dtype=torch.int64, device=self.device)
        rc = self.lib.b200fit_mask_argmax(self.ctx, _ptr(mask_counts), int(mask_counts.numel()), int(index_offset),
                                          _ptr(best), _ptr(self._ws), self._stream())
        self._check(rc, "b200fit_mask_argmax")
        return best

    def mask_bits(self, prob, params1, params2, best=None, pix=-1, model_f32=True, ncomp_pair=(1, 2)):
        """device uint32[64] bit set of the un-dilated union model mask of one pixel."""
        bits = torch.empty(64, dtype=torch.int32, device=self.device)
        rc = self.lib.b200fit_mask_bits_pair(self.ctx, C.byref(prob), int(ncomp_pair[0]), int(ncomp_pair[1]),
                                        _ptr(params1), _ptr(params2), 1 if model_f32 else 0,
                                        _ptr(best), int(pix), _ptr(bits), self._stream())
        self._check(rc, "b200fit_mask_bits")
        return bits

    def aicc_global(self, prob, spectra, params1, params2, expand=20, model_f32=True, thr=(5.0, 5.0, 5.0),
                    row_index=None, ncomp_pair=(1, 2), mask_bits=None):
        """lnk10/lnk20/lnk21, rss, N and the integer ncomp map of every pixel of ONE device's pixel set, with the
        include_nosamp rule applied over that set (single-GPU composition; dist.py composes it across ranks).
        No host synchronisation between the passes.  ``mask_bits``: see aicc_pass."""
        out = self.aicc_alloc(int(prob.npix))
        self.aicc_pass(prob, spectra, params1, params2, out, expand=expand, model_f32=model_f32, thr=thr,
                       row_index=row_index, ncomp_pair=ncomp_pair, mask_bits=mask_bits)
        best = self.mask_argmax(out["mask_counts"])
        if mask_bits is None:
            bits = self.mask_bits(prob, params1, params2, best=best, model_f32=model_f32, ncomp_pair=ncomp_pair)
        else:   # the fill mask of the include_nosamp rule is the given mask of the pixel holding the most samples
            bits = torch.where(best[0] > 0, mask_bits[best[1].clamp(0, max(int(prob.npix) - 1, 0))],
                               torch.zeros(64, dtype=torch.int32, device=self.device)).contiguous()
        self.aicc_pass(prob, spectra, params1, params2, out, fill_bits=bits, only_empty=True, expand=expand,
                       model_f32=model_f32, thr=thr, row_index=row_index, ncomp_pair=ncomp_pair, mask_bits=mask_bits)
        out["fill_bits"] = bits
        return out

    def prepare_guesses(self, raw, lo, hi, eps):
        """cubefit_simp's conditioning of the guesses on the device (b200fit_prepare_guesses): ``raw`` [P, npix] fp64
        field-major device tensor -> [npix, P] with zero velocities -> NaN and every value moved inside [lo, hi]."""
        P, npix = (int(x) for x in raw.shape)
        assert raw.dtype == torch.float64 and raw.is_contiguous() and P in (4, 8) and len(lo) == P and len(hi) == P
        out = torch.empty((npix, P), dtype=torch.float64, device=self.device)
        lo_a = (C.c_double * P)(*[float(x) for x in lo])
        hi_a = (C.c_double * P)(*[float(x) for x in hi])
        rc = self.lib.b200fit_prepare_guesses(self.ctx, npix, P, _ptr(raw), C.cast(lo_a, C.c_void_p),
                                              C.cast(hi_a, C.c_void_p), float(eps), _ptr(out), self._stream())
        self._check(rc, "b200fit_prepare_guesses")
        return out

    def pack_selection(self, fit1, fit2, sel, out=None):
        """get_best_2c_parcube assembled on the device (b200fit_pack_selection): ``fit1`` / ``fit2`` = result dicts
        of the 1- and 2-component fits over the same pixels (params, perr), ``sel`` = output of aicc_global /
        aicc_distributed.  Returns the field-major [PACK_FIELDS, npix] fp64 device tensor."""
        npix = int(fit1["params"].shape[0])
        if out is None:
            out = torch.empty((_abi.PACK_FIELDS, npix), dtype=torch.float64, device=self.device)
        assert out.is_contiguous() and tuple(out.shape) == (_abi.PACK_FIELDS, npix)
        for t, w in ((fit1["params"], 4), (fit1["perr"], 4), (fit2["params"], 8), (fit2["perr"], 8)):
            assert t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == (npix, w)
        rc = self.lib.b200fit_pack_selection(self.ctx, npix, _ptr(fit1["params"]), _ptr(fit1["perr"]),
                                             _ptr(fit2["params"]), _ptr(fit2["perr"]), _ptr(sel["lnk"]),
                                             _ptr(sel["ncomp"]), _ptr(out), npix, self._stream())
        self._check(rc, "b200fit_pack_selection")
        return out

    def rms_snr(self, nchan, spectra, planemask=None, sigma_cut=2.5, expand=20, row_index=None):
        """signals.get_rms_robust + the peak for get_snr / snr_estimate, per pixel on the GPU.  Returns device
        tensors rms0 (plain mad_std), rms (refined), peak (max over the included channels)."""
        npix = int(spectra.shape[0]) if row_index is None else int(row_index.numel())
        out = {k: torch.empty(npix, dtype=torch.float64, device=self.device) for k in ("rms0", "rms", "peak")}
        if planemask is not None:
            assert planemask.dtype == torch.uint8 and planemask.numel() == npix and planemask.is_contiguous()
        rc = self.lib.b200fit_rms_snr(self.ctx, int(nchan), npix, _ptr(spectra), spectra.shape[1], _ptr(row_index),
                                      _ptr(planemask), float(sigma_cut), int(expand), _ptr(out["rms0"]),
                                      _ptr(out["rms"]), _ptr(out["peak"]), self._stream())
        self._check(rc, "b200fit_rms_snr")
        return out

    def resid_stats(self, prob, spectra, params, expand=20, model_f32=True, row_index=None):
        npix = int(prob.npix)
        out = dict(rms=torch.empty(npix, dtype=torch.float64, device=self.device),
                   chisq_red=torch.empty(npix, dtype=torch.float64, device=self.device))
        rc = self.lib.b200fit_resid_stats(self.ctx, C.byref(prob), _ptr(spectra), spectra.shape[1], _ptr(row_index),
                                          _ptr(params), int(expand), 1 if model_f32 else 0, _ptr(out["rms"]),
                                          _ptr(out["chisq_red"]), self._stream())
        self._check(rc, "b200fit_resid_stats")
        return out

    def convolve_planes(self, cube_dev, kx=None, ky=None, k2=None, include=None, out=None):
        """Spatial convolution of every channel plane (b200fit_convolve_planes).  cube_dev [nchan, ny, nx] fp32;
        separable kernel ``kx``, ``ky`` [K] or general ``k2`` [K, K] (fp32 device tensors); ``include`` [ny, nx]
        uint8 or None.  Returns the convolved cube (device)."""
        assert cube_dev.dtype == torch.float32 and cube_dev.is_contiguous() and cube_dev.dim() == 3
        nchan, ny, nx = (int(v) for v in cube_dev.shape)
        if out is None:
            out = torch.empty_like(cube_dev)
        kern = k2 if k2 is not None else kx
        ksize = int(kern.shape[0])
        for t in (kx, ky, k2):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda)
        if include is not None:
            assert include.dtype == torch.uint8 and include.is_contiguous() and include.numel() == ny * nx
        ksum = float(k2.double().sum()) if k2 is not None else float(kx.double().sum()) * float(ky.double().sum())
        rc = self.lib.b200fit_convolve_planes(self.ctx, _ptr(cube_dev), nchan, ny, nx, _ptr(include), ksize, _ptr(kx),
                                              _ptr(ky), _ptr(k2), ksum, _ptr(out), self._stream())
        self._check(rc, "b200fit_convolve_planes")
        return out

    def resample_bilinear(self, cube_dev, out_ny, out_nx, ay, by, ax, bx):
        """Bilinear regrid of every plane: output pixel (oy, ox) samples input (ay oy + by, ax ox + bx)."""
        assert cube_dev.dtype == torch.float32 and cube_dev.is_contiguous() and cube_dev.dim() == 3
        nchan, ny, nx = (int(v) for v in cube_dev.shape)
        out = torch.empty((nchan, int(out_ny), int(out_nx)), dtype=torch.float32, device=self.device)
        rc = self.lib.b200fit_resample_bilinear(self.ctx, _ptr(cube_dev), nchan, ny, nx, _ptr(out), int(out_ny),
                                                int(out_nx), float(ay), float(by), float(ax), float(bx), self._stream())
        self._check(rc, "b200fit_resample_bilinear")
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
