"""PCube: the object MUFASA's drivers talk to where the reference uses ``pyspeckit.Cube``.

Same attribute names / shapes / dtypes the reference reads back:
  parcube[P, ny, nx], errcube[P, ny, nx] (float64, zeros where not fitted), has_fit[ny, nx] (bool),
  get_modelcube() -> [nchan, ny, nx] cached in ``_modelcube`` (NaN where no fit), cube, xarr, maskmap,
  specfit.Registry.add_fitter(...), fiteach(...) with pyspeckit's keyword names.
The per-pixel work -- what ``Cube.fiteach`` does in forked worker processes -- is one call into the CUDA library
(b200fit_gather_spectra + b200fit_fit).  Extra maps the reference does not keep are exposed too: chi2map,
statusmap, nevalmap, errmap_used.
"""
import numpy as np
import torch

from . import _abi
from .engine import get_engine
from .exceptions import FitTypeError
from .spec_models.meta_model import FITTYPE_ALIASES


class _Registry(object):
    """pyspeckit's fitter registry: MUFASA calls ``pcube.specfit.Registry.add_fitter(fittype, fitter, npars)``
    (multi_v_fit.py:119-120; UltraCube.py:851-854).  Here it maps fittype -> model id of the CUDA library."""

    def __init__(self):
        self.multifitters = {}
        self.npars = {}

    def add_fitter(self, name, function, npars, **kwargs):
        name = FITTYPE_ALIASES.get(name, name)
        if name not in _abi.MODEL_IDS:
            raise FitTypeError("\'{}\' is an invalid fittype".format(name))
        self.multifitters[name] = function
        self.npars[name] = npars


class _Specfit(object):
    def __init__(self):
        self.Registry = _Registry()


class VelocityAxis(object):
    """Uniform velocity axis in km/s, radio convention about ``refX`` (Hz).  Carries the two attributes
    register_pcube inspects (multi_v_fit.py:107-116)."""

    def __init__(self, v_kms, refX=None, velocity_convention='radio'):
        v = np.asarray(v_kms, dtype=np.float64)
        if v.ndim != 1 or v.size < 8:
            raise ValueError("spectral axis must be 1-D with at least 8 channels")
        dv = np.diff(v)
        if not np.all(np.isfinite(v)) or dv[0] == 0 or np.max(np.abs(dv - dv[0])) > 1e-6 * abs(dv[0]):
            raise ValueError("the CUDA fitter needs a uniform velocity axis")
        self.value = v
        self.refX = refX
        self.velocity_convention = velocity_convention
        self.unit = 'km/s'

    def __len__(self):
        return self.value.size

    @property
    def v0_dv_flip(self):
        """(v0, dv, flip) of the ascending axis the device works on."""
        v = self.value
        dv = (v[-1] - v[0]) / (v.size - 1)
        if dv > 0:
            return float(v[0]), float(dv), False
        return float(v[-1]), float(-dv), True


class PCube(object):
    def __init__(self, cube, xarr, header=None, unit='K', maskmap=None, device=None):
        cube = np.asarray(cube)
        if cube.ndim != 3:
            raise ValueError("cube must be [nchan, ny, nx]")
        self.cube = cube
        self.xarr = xarr if isinstance(xarr, VelocityAxis) else VelocityAxis(xarr)
        if len(self.xarr) != cube.shape[0]:
            raise ValueError("spectral axis length does not match the cube")
        self.header = header
        self.wcs = None
        self.unit = unit
        self.maskmap = maskmap
        self.specfit = _Specfit()
        self.parcube = None
        self.errcube = None
        self.has_fit = np.zeros(cube.shape[1:], dtype=bool)
        self._modelcube = None
        self.chi2map = None
        self.statusmap = None
        self.nevalmap = None
        self.errmap_used = None
        self.fittype = None
        self.npars = None
        self._device = device
        self._dev = {}            # device-resident state, shared between PCubes of one data cube
        self._deferred = None     # list collecting fit jobs when UltraCube batches several fits (see fiteach)
        self.timing = {}

    # ---------------------------------------------------------------------------------------------------
    @property
    def engine(self):
        return get_engine(self._device)

    @staticmethod
    def slab_edges(ny, nx, nslab, min_pixels):
        """Row boundaries [0, r0, ..., ny] of the ``nslab`` (>= 2) upload slabs: a short first slab (1/16 of the rows,
        at least ``min_pixels`` spectra) so that the fits start sooner, the rest in equal parts."""
        r0 = max(ny // 16, -(-min_pixels // nx))
        return [0] + [r0 + k * (ny - r0) // (nslab - 1) for k in range(nslab)]

    def rows_on_device(self, wait=True, max_slabs=None):
        """The cube as pixel-major rows [ny*nx, stride] fp32 on the GPU (NaN padded, velocity ascending).

        Uploaded ONCE per data cube and shared between the PCubes of one UltraCube (``share_device``); the reference
        instead re-reads the FITS file for every ncomp (UltraCube.py:191).  A large cube in pinned memory goes up as
        up to four row slabs on a second stream (``b200fit_upload_slab`` + ``b200fit_gather_spectra`` per slab), each
        with an event: fit jobs that only touch the first slab start while the others are still crossing PCIe
        (``slab_ready``, used by Engine.fit_batch).  ``wait=True`` orders the current stream after the whole upload
        -- what every consumer other than the batched fit wants.  ``max_slabs`` = how many slabs must have
        been QUEUED when this call returns (None: all): host-to-device copies are served in the order they were queued, whatever
        their stream, so a caller with small uploads of its own (guesses) queues one slab, then those, then the rest
        (dist.SelectionJob) -- queued behind the whole cube they would arrive after it, and so would the first fit."""
        if self._dev.get("rows") is None:
            nchan, ny, nx = self.cube.shape
            eng = self.engine
            host = self.cube if self.cube.dtype == np.float32 else self.cube.astype(np.float32)
            t = torch.from_numpy(np.ascontiguousarray(host).reshape(nchan, ny * nx))
            _, _, flip = self.xarr.v0_dv_flip
            # slabs of at least SPLIT_MIN_PIXELS spectra, at most 4 (two fits x 4 slabs fill a batch of the library)
            nslab = 1
            if t.is_pinned():
                nslab = int(max(1, min(4, ny, (ny * nx) // eng.SPLIT_MIN_PIXELS)))
            if nslab == 1:
                staging = t.to(eng.device, non_blocking=True)       # async when the cube lives in pinned memory
                self._dev["rows"] = eng.gather(staging, None, flip=flip)
                del staging
            else:
                rows = torch.empty((ny * nx, eng.nchan_stride(nchan)), dtype=torch.float32, device=eng.device)
                up = eng.side_stream("h2d_cube")
                up.wait_stream(torch.cuda.current_stream(eng.device))   # the allocation above is ordered before
                bounds, events = [], []

                edges = self.slab_edges(ny, nx, nslab, eng.SPLIT_MIN_PIXELS)

                def queue_slabs():
                    for k in range(nslab):
                        first, last = edges[k] * nx, edges[k + 1] * nx
                        with torch.cuda.stream(up):
                            staging = torch.empty((nchan, last - first), dtype=torch.float32, device=eng.device)
                            # in pieces of ~256 MB (one 2-D copy each; small pieces only cost host time now that
                            # nothing has to slip in between them -- the callers queue their own uploads in order)
                            step = max(1, (256 << 20) // (4 * (last - first)))
                            for c0 in range(0, nchan, step):
                                c1 = min(nchan, c0 + step)
                                eng.upload_slab(t[c0:c1], first, last - first, staging[c0:c1])
                            eng.gather(staging, None, flip=flip, out=rows[first:last])
                            ev = torch.cuda.Event()
                            ev.record(up)
                            del staging
                        bounds.append(last)
                        events.append(ev)
                        yield

                self._dev["rows"] = rows
                self._dev["slab_ready"] = (bounds, events)
                self._dev["slab_queue"] = queue_slabs()
                self._dev["host_cube"] = t                      # keeps the pinned source alive until the copies ran
            self._dev["h2d_bytes"] = t.numel() * 4
        pending = self._dev.get("slab_queue")
        if pending is not None:                 # ``max_slabs`` = how many slabs must have been queued in total
            while max_slabs is None or wait or self._dev.get("slabs_queued", 0) < max_slabs:
                try:
                    next(pending)
                except StopIteration:
                    self._dev["slab_queue"] = None
                    break
                self._dev["slabs_queued"] = self._dev.get("slabs_queued", 0) + 1
        if wait and self._dev.get("slab_ready") is not None:
            torch.cuda.current_stream(self.engine.device).wait_event(self._dev["slab_ready"][1][-1])
        return self._dev["rows"]

    def footprint(self):
        """[ny, nx] bool: pixels with at least one finite channel (np.any(np.isfinite(cube), axis=0))."""
        if self._dev.get("footprint") is None:
            rows = self.rows_on_device()
            nchan = self.cube.shape[0]
            fp = torch.isfinite(rows[:, :nchan]).any(dim=1)
            self._dev["footprint"] = fp.cpu().numpy().reshape(self.cube.shape[1:])
        return self._dev["footprint"]

    def rms_robust(self, planemask=None, sigma_cut=2.5, expand=20):
        """signals.get_rms_robust / the peak for get_snr, per pixel on the GPU (b200fit_rms_snr).  ``planemask``
        [ny, nx] bool = pixels kept by trim_cube_edge (None: all).  Returns dict of [ny, nx] float64 maps: rms0
        (= mad_std(cube, axis=0)), rms (refined), peak.  Cached per (planemask, sigma_cut, expand)."""
        ny, nx = self.cube.shape[1:]
        key = ("rms", None if planemask is None else np.asarray(planemask, bool).tobytes(), float(sigma_cut), int(expand))
        cache = self._dev.setdefault("rms_cache", {})
        if key not in cache:
            eng = self.engine
            rows = self.rows_on_device()
            pm = None
            if planemask is not None:
                pm = torch.from_numpy(np.ascontiguousarray(planemask, dtype=np.uint8).ravel()).to(eng.device)
            out = eng.rms_snr(self.cube.shape[0], rows, planemask=pm, sigma_cut=sigma_cut, expand=expand)
            packed = torch.stack([out["rms0"], out["rms"], out["peak"]]).cpu().numpy()
            cache.clear()
            cache[key] = dict(rms0=packed[0].reshape(ny, nx), rms=packed[1].reshape(ny, nx),
                              peak=packed[2].reshape(ny, nx))
        return cache[key]

    def share_device(self, other):
        """Use the device-resident rows of another PCube built on the same data cube."""
        if other.cube is not self.cube and other.cube.shape != self.cube.shape:
            raise ValueError("cannot share device data between different cubes")
        self._dev = other._dev

    def release_device(self):
        self._dev.clear()

    def copy(self):
        new = PCube(self.cube, self.xarr, header=self.header, unit=self.unit, maskmap=self.maskmap,
                    device=self._device)
        new._dev = self._dev
        for k in ("parcube", "errcube", "has_fit", "_modelcube", "chi2map", "statusmap", "nevalmap", "errmap_used"):
            v = getattr(self, k)
            setattr(new, k, None if v is None else v.copy())
        new.fittype, new.npars = self.fittype, self.npars
        new.specfit = self.specfit
        return new

    def _problem(self, fittype, ncomp, npix, minpars, maxpars, signal_cut=0.0, maxiter=200):
        v0, dv, _ = self.xarr.v0_dv_flip
        nu_ref = 0.0 if self.xarr.refX is None else float(self.xarr.refX) / 1e9
        return _abi.make_problem(fittype, ncomp, self.cube.shape[0], npix, v0, dv, minpars, maxpars,
                                 nu_ref_ghz=nu_ref, max_iter=maxiter, signal_cut=signal_cut)

    # ---------------------------------------------------------------------------------------------------
    def fiteach(self, fittype='nh3_multi_v', guesses=None, limitedmin=None, limitedmax=None, minpars=None,
                maxpars=None, multicore=1, maskmap=None, start_from_point=None, use_neighbor_as_guess=False,
                guesses_are_moments=False, signal_cut=3, errmap=None, verbose_level=0, verbose=False,
                maxiter=200, **kwargs):
        """pyspeckit ``Cube.fiteach`` for the call MUFASA makes (multi_v_fit.py:374-393, :192-206).

        Every valid pixel (maskmap & finite guesses & a finite channel) is fitted on the GPU in one launch.
        ``multicore``, ``start_from_point`` and the verbosity switches are accepted and ignored: pixel order does
        not matter because use_neighbor_as_guess=False (multi_v_fit.py:381).  ``signal_cut`` keeps pyspeckit's
        default of 3; the MUFASA drivers pass 0."""
        fittype = FITTYPE_ALIASES.get(fittype, fittype)
        if fittype not in _abi.MODEL_IDS:
            raise FitTypeError("\'{}\' is an invalid fittype".format(fittype))
        if use_neighbor_as_guess or guesses_are_moments:
            raise NotImplementedError("MUFASA always passes use_neighbor_as_guess=False, guesses_are_moments=False")
        if errmap is not None:
            raise NotImplementedError("MUFASA never forwards an errmap to fiteach (multi_v_fit.py:234-238)")
        guesses = np.asarray(guesses, dtype=np.float64)
        nchan, ny, nx = self.cube.shape
        if guesses.ndim != 3 or guesses.shape[1:] != (ny, nx) or guesses.shape[0] % 4:
            raise ValueError("guesses must be [4*ncomp, ny, nx]")
        P = guesses.shape[0]
        ncomp = P // 4
        if ncomp not in (1, 2):
            raise NotImplementedError("the CUDA fitter covers 1- and 2-component models")
        for lim in (limitedmin, limitedmax):
            if lim is not None and not all(lim):
                raise NotImplementedError("all parameter limits are active in MUFASA's calls")
        if minpars is None or maxpars is None or len(minpars) != P or len(maxpars) != P:
            raise ValueError("minpars/maxpars must have one entry per parameter")

        ok = np.all(np.isfinite(guesses), axis=0)
        if maskmap is not None:
            ok &= np.asarray(maskmap, dtype=bool)
        if self.maskmap is not None:
            ok &= np.asarray(self.maskmap, dtype=bool)
        if start_from_point is not None:
            sx, sy = int(start_from_point[0]), int(start_from_point[1])
            if not (0 <= sx < nx and 0 <= sy < ny) or not ok[sy, sx]:
                raise ValueError("The starting fit position is not among the valid pixels")   # pyspeckit contract
        pix = np.flatnonzero(ok.ravel())
        npix = pix.size

        self.parcube = np.zeros((P, ny, nx), dtype=np.float64)
        self.errcube = np.zeros((P, ny, nx), dtype=np.float64)
        self.has_fit = np.zeros((ny, nx), dtype=bool)
        self.chi2map = np.zeros((ny, nx), dtype=np.float64)
        self.statusmap = np.zeros((ny, nx), dtype=np.int32)
        self.nevalmap = np.zeros((ny, nx), dtype=np.int32)
        self.errmap_used = np.full((ny, nx), np.nan)
        self._modelcube = None
        self.fittype, self.npars = fittype, P
        self.last_minpars, self.last_maxpars = list(minpars), list(maxpars)
        self.last_guesses, self.last_maskmap = guesses, ok
        if npix == 0:
            return

        eng = self.engine
        dev = eng.device
        prob = self._problem(fittype, ncomp, npix, list(minpars), list(maxpars), signal_cut=float(signal_cut or 0.0),
                             maxiter=maxiter)
        rows = self.rows_on_device(wait=False, max_slabs=1)   # the other slabs follow the guesses (copies are FIFO)
        # guesses (and the pixel list) go up through pinned memory on a second stream: a copy from pageable memory
        # on the main stream would make the host wait for the cube upload queued ahead of it
        main = torch.cuda.current_stream(dev)
        up = eng.side_stream("h2d")
        # [P, npix] on the host (a plain copy when every pixel is fitted), transposed to [npix, P] on the device
        g_stage = eng.pinned_stage(("guesses", P, npix))
        gflat = guesses.reshape(P, ny * nx)
        np.copyto(g_stage.numpy(), gflat if npix == ny * nx else gflat[:, pix])
        i_stage = None
        if npix != ny * nx:
            # one buffer per model size: the 1- and 2-component jobs of a batched fit may select different pixels, and
            # the first job's (asynchronous) copy must not read the second job's list
            i_stage = eng.pinned_stage(("pixels%d" % P, npix), dtype=torch.int64)
            np.copyto(i_stage.numpy(), pix)
        with torch.cuda.stream(up):
            g_dev = g_stage.to(dev, non_blocking=True).t().contiguous()
            idx_dev = None if i_stage is None else i_stage.to(dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(up)
        g_dev.record_stream(main)
        if idx_dev is not None:
            idx_dev.record_stream(main)
        if self._deferred is None:
            self.rows_on_device(wait=False)     # (a batched fit queues the rest once every job's guesses are on their way)
        job = dict(prob=prob, spectra=rows, guesses=g_dev, row_index=idx_dev, pcube=self, pix=pix, ready=ready,
                   slab_ready=self._dev.get("slab_ready"))
        if self._deferred is not None:
            # UltraCube.fit_cube over several ncomp: the fits are launched together (Engine.fit_batch) so that the
            # latency-bound late rounds of one overlap the other; results are filled in by finish_fit()
            self._deferred.append(job)
            return
        out = eng.fit_batch([job])[0]
        self.finish_fit(job, out)

    def start_fit_readback(self, job, out):
        """Queue the D2H copy of a fit's results on the current stream (no host synchronisation): one device-side
        transpose to field-major [2P + 4, npix], one copy into a pinned buffer."""
        packed_d = torch.cat([out["params"], out["perr"], out["chi2"][:, None], out["err_used"][:, None],
                              out["status"].to(torch.float64)[:, None],
                              out["counts"][:, 0].to(torch.float64)[:, None]], dim=1).t().contiguous()
        stage = self.engine.pinned_stage(tuple(packed_d.shape))   # pinned staging buffers live with the engine
        stage.copy_(packed_d, non_blocking=True)
        job["host_stage"] = stage
        return stage

    def finish_fit(self, job, out):
        """Copy the results of a (possibly batched) fit back into parcube / errcube / has_fit ..."""
        pix = job["pix"]
        P = self.npars
        ny, nx = self.has_fit.shape
        # readback (start_fit_readback), then contiguous row copies into the maps
        stage = job.get("host_stage")
        if stage is None:
            stage = self.start_fit_readback(job, out)
        torch.cuda.current_stream(out["params"].device).synchronize()
        packed = stage.numpy()
        flat = lambda a: a.reshape(a.shape[0], ny * nx) if a.ndim == 3 else a.reshape(ny * nx)
        if pix.size == ny * nx:
            pix = slice(None)           # every pixel fitted: plain copies instead of fancy-index scatters
        flat(self.parcube)[:, pix] = packed[:P]
        flat(self.errcube)[:, pix] = packed[P:2 * P]
        flat(self.chi2map)[pix] = packed[2 * P]
        flat(self.errmap_used)[pix] = packed[2 * P + 1]
        st = packed[2 * P + 2].astype(np.int32)
        flat(self.statusmap)[pix] = st
        flat(self.nevalmap)[pix] = packed[2 * P + 3].astype(np.int32)
        flat(self.has_fit)[pix] = st > 0
        if not self.has_fit.any():
            # pyspeckit: "The first fitted pixel did not yield a fit" -> AssertionError, which cubefit_gen turns
            # into a retry / StartFitError (multi_v_fit.py:388-403)
            raise AssertionError("No pixel yielded a fit. Please check the guesses, limits and mask.")

    # ---------------------------------------------------------------------------------------------------
    def load_model_fit(self, fitsfilename, npars, fittype=None, **kwargs):
        """pyspeckit ``Cube.load_model_fit`` for MUFASA's products (UltraCube.py:829-858): planes [0, npars) are the
        parameters, [npars, 2 npars) their errors (multi_v_fit.save_pcube)."""
        from .utils import fits_io
        data, hdr = fits_io.read(fitsfilename)
        if data.ndim != 3 or data.shape[0] < 2 * npars or data.shape[1:] != self.cube.shape[1:]:
            raise ValueError("{}: expected [>= {}, {}, {}] planes".format(fitsfilename, 2 * npars, *self.cube.shape[1:]))
        self.parcube = np.array(data[:npars], dtype=np.float64)
        self.errcube = np.array(data[npars:2 * npars], dtype=np.float64)
        with np.errstate(invalid="ignore"):
            self.has_fit = np.any(self.parcube != 0, axis=0) & np.all(np.isfinite(self.parcube), axis=0)
        self.npars = npars
        self.fittype = FITTYPE_ALIASES.get(fittype, fittype) if fittype is not None else self.fittype
        self._modelcube = None
        self.fit_header = hdr
        return self

    def model_of_pixels(self, pix):
        """[nchan, len(pix)] model spectra of the flat pixel indices ``pix`` (NaN where a pixel has no fit), in the data
        cube's channel order and dtype -- what get_modelcube would hold in those columns, without the cube."""
        P, ny, nx = self.parcube.shape
        nchan = self.cube.shape[0]
        pix = np.asarray(pix, dtype=np.int64)
        dtype = self.cube.dtype if self.cube.dtype.kind == 'f' else np.float64
        out = np.full((nchan, pix.size), np.nan, dtype=dtype)
        par = self.parcube.reshape(P, ny * nx)[:, pix]
        valid = np.any(par != 0, axis=0) & np.all(np.isfinite(par), axis=0)
        if valid.any():
            eng = self.engine
            prob = self._problem(self.fittype, P // 4, int(valid.sum()), [-1e30] * P, [1e30] * P)
            m = eng.model(prob, torch.from_numpy(np.ascontiguousarray(par[:, valid].T)).to(eng.device))[:, :nchan]
            _, _, flip = self.xarr.v0_dv_flip
            if flip:
                m = torch.flip(m, dims=[1])
            out[:, valid] = m.to(torch.float32 if dtype == np.float32 else torch.float64).cpu().numpy().T
        return out

    def get_modelcube(self, update=False, multicore=1, mask=None):
        """Model cube from parcube (pyspeckit get_modelcube: NaN where there is no fit), dtype of the data cube,
        cached in ``_modelcube``; callers may mutate the returned array (master_fitter.py:2516-2520)."""
        if self._modelcube is not None and not update:
            return self._modelcube
        if self.parcube is None:
            raise AttributeError("no fit has been performed")
        P, ny, nx = self.parcube.shape
        nchan = self.cube.shape[0]
        valid = np.any(self.parcube != 0, axis=0) & np.all(np.isfinite(self.parcube), axis=0)
        if mask is not None:
            valid &= np.asarray(mask, dtype=bool)
        out = np.full((nchan, ny, nx), np.nan, dtype=self.cube.dtype if self.cube.dtype.kind == 'f' else np.float64)
        pix = np.flatnonzero(valid.ravel())
        if pix.size:
            eng = self.engine
            lo = [-1e30] * P
            hi = [1e30] * P
            prob = self._problem(self.fittype, P // 4, pix.size, lo, hi)
            par = np.ascontiguousarray(self.parcube.reshape(P, ny * nx)[:, pix].T)
            m = eng.model(prob, torch.from_numpy(par).to(eng.device))[:, :nchan]
            _, _, flip = self.xarr.v0_dv_flip
            if flip:
                m = torch.flip(m, dims=[1])
            out.reshape(nchan, ny * nx)[:, pix] = m.to(torch.float32 if out.dtype == np.float32 else torch.float64
                                                      ).cpu().numpy().T
        self._modelcube = out
        return out
