"""UltraCube: host-side mirror of mufasa/UltraCube.py for the B200 fitter.

Keeps the reference's method names and result conventions
  * UltraCube.fit_cube                      UltraCube.py:272-326 (+ module fit_cube :777-801)
  * has_fit / include_model_mask / reset_model_mask           :328-368
  * get_rss / get_AICc / get_AICc_likelihood / get_all_lnk_maps / get_best_2c_parcube   :432-498, :1120-1379
  * get_residual / get_rms / get_reduced_chisq                :403-471, :1478-1565, :1701-1800
while every per-pixel computation is a kernel of libb200fit.so.  Differences that come from living on a GPU:
  * the cube is uploaded once and kept resident as pixel-major rows (the reference re-reads the FITS file for every
    ncomp, UltraCube.py:191) -- both fits and the AICc pass stream the same device buffer;
  * ``master_model_mask`` (a [nchan, ny, nx] bool cube, 1 GB at 1024x1024x1000) is materialised only when a caller
    asks for it; the AICc kernels rebuild the per-pixel mask from the fitted parameters on the fly.
"""
import gc
import logging
import os
import weakref
from collections.abc import Iterable

import numpy as np
import torch

from . import multi_v_fit as mvf
from .cube import SpecCube
from .exceptions import StartFitError
from .pcube import PCube
from .spec_models.meta_model import MetaModel

logger = logging.getLogger(__name__)


class UltraCube(object):
    def __init__(self, cubefile=None, cube=None, fittype=None, snr_min=None, rmsfile=None, cnv_factor=2,
                 n_cores=True, device=None, share_device_with=None):
        self.pcubes = {}
        self.residual_cubes = {}
        self.rms_maps = {}
        self.Tpeak_maps = {}
        self.chisq_maps = {}
        self.rchisq_maps = {}
        self.rss_maps = {}
        self.NSamp_maps = {}
        self.AICc_maps = {}
        self.lnk_maps = {}
        self.ncomp_map = None
        self._ncomp_thr = None
        self._master_model_mask = None
        self._aicc_eager = None
        self._mask_ncomps = []            # which fits feed the lazily built master_model_mask
        self.snr_min = 0.0
        self.cnv_factor = cnv_factor
        self.n_cores = mvf.validate_n_cores(n_cores)
        self.fittype = fittype
        self.meta_model = None
        self.device = device
        self.cubefile = cubefile
        if cubefile is not None and cube is None:
            cube = SpecCube.read(cubefile)
        if not isinstance(cube, SpecCube):
            raise TypeError("cube must be a mufasa_b200.cube.SpecCube")
        self.cube = cube
        if fittype is not None:
            self.meta_model = MetaModel(fittype=fittype, ncomp=1)
        if snr_min is not None:
            self.snr_min = snr_min
        if rmsfile is not None:
            self.rmsfile = rmsfile
        self._ref_pcube = None
        self.dist = None                  # (process group, global index of this slab's first pixel) when partitioned
        if share_device_with is not None:
            # a second UltraCube on the SAME data (replace_bad_pix builds one per refit, master_fitter.py:1437):
            # reuse the rows already resident on the GPU instead of uploading the cube again
            other = share_device_with._ref_pcube or share_device_with.load_pcube()
            self._ref_pcube = self.load_pcube(pcube_ref=other)

    # ---- pcube handling ---------------------------------------------------------------------------------
    def load_pcube(self, pcube_ref=None):
        """UltraCube.py:173-214: a fresh fit container on the same data.  The device copy of the cube is shared."""
        p = PCube(self.cube._data, self.cube.xarr(), header=self.cube.header, unit=self.cube.unit,
                  device=self.device)
        ref = pcube_ref if pcube_ref is not None else self._ref_pcube
        if ref is not None:
            p.share_device(ref)
        if self._ref_pcube is None:
            self._ref_pcube = p
        return p

    def convolve_cube(self, savename=None, factor=None, edgetrim_width=5):
        """Smooth the cube to ``factor`` x its beam and regrid it onto ``factor`` x larger pixels (UltraCube.py:218-
        249 -> convolve_tools.convolve_sky_byfactor); the result is kept in ``self.cube_cnv``."""
        from . import convolve_tools as cnvtool
        if factor is None:
            factor = self.cnv_factor
        self.cube_cnv = cnvtool.convolve_sky_byfactor(self.cube, factor, savename, edgetrim_width=edgetrim_width,
                                                      device=self.device)

    def fit_cube(self, ncomp, simpfit=False, concurrent=True, eager_aicc=True, **kwargs):
        """UltraCube.py:272-326.  With several ``ncomp`` the reference fits them one after the other; here the host
        drivers still run in that order, but the GPU fits are launched together (``concurrent``; Engine.fit_batch)
        so the slow tail of one fit overlaps the other.  Results are identical to sequential fits."""
        if 'multicore' not in kwargs:
            kwargs['multicore'] = self.n_cores
        if 'snr_min' not in kwargs:
            kwargs['snr_min'] = self.snr_min
        if not isinstance(ncomp, Iterable):
            ncomp = [ncomp]
        ncomp = list(ncomp)
        batch = [] if (concurrent and len(ncomp) > 1) else None
        pending = []
        eager = None
        try:
            for nc in ncomp:
                if self.pcubes:
                    _, pcube_ref = next(iter(self.pcubes.items()))
                    self.pcubes[str(nc)] = self.load_pcube(pcube_ref=pcube_ref)
                else:
                    self.pcubes[str(nc)] = self.load_pcube()
                self.pcubes[str(nc)]._deferred = batch
                kw = dict(kwargs)
                if isinstance(kw.get('guesses'), dict):          # per-ncomp guesses for a batched call
                    kw['guesses'] = kw['guesses'].get(nc)
                if simpfit and batch is not None:
                    # first half of every driver (host work on the guesses) while the cube is still uploading; the
                    # halves that need the device follow below
                    steps = mvf.cubefit_simp_steps(self.cube, self.pcubes[str(nc)], fittype=self.fittype, ncomp=nc,
                                                   **kw)
                    next(steps)
                    pending.append((nc, steps))
                    continue
                self.pcubes[str(nc)] = fit_cube(self.cube, self.pcubes[str(nc)], fittype=self.fittype,
                                                simpfit=simpfit, ncomp=nc, **kw)
            for nc, steps in pending:
                for _ in steps:
                    pass
            if batch:
                eng = batch[0]["pcube"].engine
                for job in batch:                               # every job's guesses are queued: now the rest of the cube
                    job["pcube"].rows_on_device(wait=False)
                outs = eng.fit_batch(batch)
                # the model selection needs nothing from the host: queue it right behind the fits, on the parameters
                # as they lie on the device, and bring the fit results back on a second stream meanwhile
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(eng.device))
                if eager_aicc:
                    eager = self._launch_eager_aicc(batch, outs, eng)
                back = eng.side_stream("d2h")
                with torch.cuda.stream(back):
                    back.wait_event(done)
                    for job, out in zip(batch, outs):           # both copies in flight before the first host wait
                        job["pcube"].start_fit_readback(job, out)
                    for job, out in zip(batch, outs):
                        try:
                            job["pcube"].finish_fit(job, out)
                        except AssertionError as e:
                            # the reference's contract at this point (multi_v_fit.py:388-403, :456-478): cubefit_gen
                            # turns "no pixel yielded a fit" into StartFitError (every pixel was tried here, so the
                            # retry from another start pixel has nothing left to do); cubefit_simp lets it through
                            if simpfit:
                                raise
                            raise StartFitError("The first two attempts to fit a pixel did not yield a fit. Message "
                                                "from the fitter: " + str(e))
                for job in batch:                                # writers the drivers deferred until the maps exist
                    save = getattr(job["pcube"], "_save_after_fit", None)
                    if save is not None:
                        mvf.save_pcube(job["pcube"], save[0], ncomp=save[1])
                        job["pcube"]._save_after_fit = None
        finally:
            for nc in ncomp:
                if str(nc) in self.pcubes:
                    self.pcubes[str(nc)]._deferred = None
        for nc in ncomp:
            p = self.pcubes[str(nc)]
            if p.has_fit.sum() > 0 and p.parcube is not None:
                self._include_fit_in_mask(nc)
            self._invalidate_stats()
        self._aicc_eager = eager

    def _launch_eager_aicc(self, batch, outs, eng, expand=20, thr=(5.0, 5.0, 5.0)):
        """AICc 0/1/2 selection of a batched 1- + 2-component fit, launched asynchronously on the device-resident fit
        output (the same arrays ``_params_dev`` would rebuild from parcube: zeros where there is no fit).  Returns
        (key, device result dict) for ``_run_aicc`` or None when the batch is not that pair."""
        by_nc = {int(job["prob"].ncomp): (job, out) for job, out in zip(batch, outs)}
        if set(by_nc) != {1, 2} or self.dist is not None:
            return None
        ref = batch[0]["pcube"]
        ny, nx = self.cube.shape[1:]
        pars = {}
        for nc, (job, out) in by_nc.items():
            if job["row_index"] is None:
                pars[nc] = out["params"]
            else:
                full = torch.zeros((ny * nx, 4 * nc), dtype=torch.float64, device=eng.device)
                full[job["row_index"]] = out["params"]
                pars[nc] = full
        prob = ref._problem(self.fittype, 1, ny * nx, [-1e30] * 4, [1e30] * 4)
        model_f32 = self.cube._data.dtype == np.float32
        out = eng.aicc_global(prob, ref.rows_on_device(), pars[1], pars[2], expand=expand, model_f32=model_f32, thr=thr)
        out["host_stage"] = self._aicc_to_host(out, eng)
        eng._aicc_stage_owner = weakref.ref(self)   # the pinned buffer is shared: a later selection takes it over
        out["host_ready"] = torch.cuda.Event()
        out["host_ready"].record(torch.cuda.current_stream(eng.device))
        return (int(expand), tuple(float(t) for t in thr)), out, pars

    # ---- on-disk contract ------------------------------------------------------------------
    def save_fit(self, savename, ncomp, header_note=None):
        """UltraCube.py:371-376 -> save_fit :805-826 -> multi_v_fit.save_pcube."""
        p = self.pcubes.get(str(ncomp))
        if p is not None and p.parcube is not None:
            mvf.save_pcube(p, savename, ncomp, header_note=header_note)
        else:
            logger.warning("no fit was performed and thus no file will be saved")

    def load_model_fit(self, filename, ncomp, calc_model=True, multicore=None):
        """UltraCube.py:379-387 -> load_model_fit :829-858: resume from a saved ``{root}_{n}vcomp.fits``."""
        p = self.load_pcube(pcube_ref=self._ref_pcube)
        mvf.register_pcube(p, MetaModel(fittype=self.fittype, ncomp=ncomp))
        p.load_model_fit(filename, npars=4 * ncomp, fittype=self.fittype)
        self.pcubes[str(ncomp)] = p
        if calc_model and p.has_fit.any():
            self._include_fit_in_mask(ncomp)
        self._invalidate_stats()

    def _invalidate_stats(self):
        for d in (self.rss_maps, self.NSamp_maps, self.AICc_maps, self.lnk_maps, self.residual_cubes, self.rms_maps,
                  self.rchisq_maps):
            d.clear()
        self.ncomp_map = None
        self._ncomp_thr = None
        self._aicc_eager = None        # a selection queued behind an earlier fit no longer describes the fits

    def has_fit(self, ncomp):
        """UltraCube.py:328-348."""
        parcube = self.pcubes[str(ncomp)].parcube
        mask = np.any(parcube != 0, axis=0)
        return mask & np.any(np.isfinite(parcube), axis=0)

    # ---- model mask (lazy) --------------------------------------------------------------------------------
    def _include_fit_in_mask(self, nc):
        if nc not in self._mask_ncomps:
            self._mask_ncomps.append(nc)
        self._master_model_mask = None

    @property
    def master_model_mask(self):
        """(model_nc > 0) OR-ed over the fits performed so far (UltraCube.py:351-368), built on demand."""
        if self._master_model_mask is None and self._mask_ncomps:
            m = None
            for nc in self._mask_ncomps:
                with np.errstate(invalid="ignore"):
                    mm = self.pcubes[str(nc)].get_modelcube() > 0
                m = mm if m is None else np.logical_or(m, mm)
            self._master_model_mask = m
        return self._master_model_mask

    @master_model_mask.setter
    def master_model_mask(self, value):
        self._master_model_mask = value
        if value is None:
            self._mask_ncomps = []

    def include_model_mask(self, mask):
        cur = self.master_model_mask
        self._master_model_mask = mask if cur is None else np.logical_or(cur, mask)

    def reset_model_mask(self, ncomps, multicore=True):
        self._master_model_mask = None
        self._mask_ncomps = [nc for nc in ncomps if nc > 0 and str(nc) in self.pcubes
                             and self.pcubes[str(nc)].parcube is not None]

    # ---- statistics: one GPU pass gives rss_0/1/2, N, AICc_0/1/2 and the three lnk maps --------------------------
    def _params_dev(self, nc, eng):
        ny, nx = self.cube.shape[1:]
        P = 4 * nc
        if str(nc) in self.pcubes and self.pcubes[str(nc)].parcube is not None:
            # field-major as it lies on the host; transposed to [npix, P] (and non-finite values zeroed) on the device
            # -- the host-side transpose was 2.4 ms per call at 512 x 512, four calls per masked refit
            par = torch.from_numpy(np.ascontiguousarray(self.pcubes[str(nc)].parcube, dtype=np.float64)
                                   ).to(eng.device, non_blocking=True).reshape(P, ny * nx)
            return torch.nan_to_num(par, nan=0.0, posinf=0.0, neginf=0.0).t().contiguous()
        return torch.zeros((ny * nx, P), dtype=torch.float64, device=eng.device)

    @staticmethod
    def _aicc_to_host(out, eng):
        packed_d = torch.cat([out["rss"], out["nsamp"][:, None], out["lnk"], out["ncomp"].to(torch.float64)[:, None]],
                             dim=1).t().contiguous()
        stage = eng.pinned_stage(("aicc",) + tuple(packed_d.shape))
        stage.copy_(packed_d, non_blocking=True)
        return stage

    def _select_ncomp(self, thr):
        """The integer 0/1/2 map for the thresholds ``thr`` = (lnk10, lnk20, lnk21), from the likelihood maps as the
        reference selects (UltraCube.py:1350-1367): 2 where lnk21 > t21 and lnk20 > t20, else 1, and 0 where not
        lnk10 > t10 (comparisons with NaN are False)."""
        with np.errstate(invalid="ignore"):
            lnk10 = -1.0 * (self.AICc_maps["1"] - self.AICc_maps["0"]) / 2.0
            lnk20 = -1.0 * (self.AICc_maps["2"] - self.AICc_maps["0"]) / 2.0
            lnk21 = -1.0 * (self.AICc_maps["2"] - self.AICc_maps["1"]) / 2.0
            two = np.logical_and(lnk21 > thr[2], lnk20 > thr[1])
            nc = np.where(two, 2, 1).astype(np.int8)
            nc[~(lnk10 > thr[0])] = 0
        self.ncomp_map = nc
        self._ncomp_thr = tuple(float(t) for t in thr)

    def _run_aicc(self, expand=20, thr=None):
        """One AICc pass per (fits, expand): rss_0/1/2, N, AICc_0/1/2, the lnk maps and the selection map.  The
        statistics do not depend on the selection thresholds, so a call with other thresholds reuses the cached maps
        and only redoes the selection (``thr`` None: whatever selection is there stays)."""
        if self.lnk_maps.get("_key") == int(expand):
            if thr is not None and tuple(float(t) for t in thr) != self._ncomp_thr:
                self._select_ncomp(thr)
            return
        thr = (5.0, 5.0, 5.0) if thr is None else tuple(float(t) for t in thr)
        key = (int(expand), thr)
        ref = self._ref_pcube if self._ref_pcube is not None else self.load_pcube()
        eng = ref.engine
        ny, nx = self.cube.shape[1:]
        eager, self._aicc_eager = getattr(self, "_aicc_eager", None), None
        kernel_thr = thr
        if eager is not None and eager[0][0] == int(expand):
            out = eager[1]              # already computed behind the batched fit (fit_cube)
            kernel_thr = eager[0][1]    # ... with these thresholds; the selection is redone below if others are asked
        else:
            out = None
        rows = ref.rows_on_device()
        prob = ref._problem(self.fittype, 1, ny * nx, [-1e30] * 4, [1e30] * 4)
        model_f32 = self.cube._data.dtype == np.float32
        if out is not None:
            pass
        elif self.dist is None:
            p1 = self._params_dev(1, eng)
            p2 = self._params_dev(2, eng)
            out = eng.aicc_global(prob, rows, p1, p2, expand=expand, model_f32=model_f32, thr=thr)
        else:       # row slab of a partitioned cube: the include_nosamp fill mask is global (dist.py)
            from . import dist as D
            p1 = self._params_dev(1, eng)
            p2 = self._params_dev(2, eng)
            out = D.aicc_distributed(eng, prob, rows, p1, p2, first_pixel=self.dist[1], expand=expand,
                                     model_f32=model_f32, thr=thr, group=self.dist[0])
        # one field-major D2H copy through a pinned buffer: rss0..2, nsamp, lnk10/20/21, ncomp
        owner = getattr(eng, "_aicc_stage_owner", None)
        if out.get("host_stage") is not None and owner is not None and owner() is self:
            out["host_ready"].synchronize()         # the eager path started the copy right behind the kernels
            packed = out["host_stage"].numpy()
            eng._aicc_stage_owner = None
        else:
            eng._aicc_stage_owner = None
            packed = self._aicc_to_host(out, eng)
            torch.cuda.current_stream(eng.device).synchronize()
            packed = packed.numpy()
        rss = packed[0:3].reshape(3, ny, nx)
        nsamp = packed[3].reshape(ny, nx)
        lnk = packed[4:7].reshape(3, ny, nx)
        with np.errstate(all="ignore"):
            for nc in (0, 1, 2):
                self.rss_maps[str(nc)] = rss[nc].copy()
                ns = np.where(np.isnan(rss[nc]), np.nan, nsamp)
                self.NSamp_maps[str(nc)] = ns
                p = 4 * nc
                self.AICc_maps[str(nc)] = ns * np.log(self.rss_maps[str(nc)] / ns) + 2 * p + \
                    (2 * p * (p + 1)) / (ns - p - 1)
        self.lnk_maps = {"_key": int(expand), "10": lnk[0].copy(), "20": lnk[1].copy(), "21": lnk[2].copy()}
        self.ncomp_map = packed[7].astype(np.int8).reshape(ny, nx)     # the kernel's selection
        self._ncomp_thr = kernel_thr
        if kernel_thr != thr:
            self._select_ncomp(thr)

    def _masked_stats(self, mask, expand):
        """rss_0/1/2 and N over the caller's spectral mask ``mask`` [nchan, ny, nx] bool (dilated by ``expand``, empty
        pixels filled by the include_nosamp rule, UltraCube.py:1423-1451): one b200fit_aicc_pair_masked pass.
        Returns (rss [3, ny, nx], nsamp [ny, nx])."""
        ref = self._ref_pcube if self._ref_pcube is not None else self.load_pcube()
        eng = ref.engine
        nchan, ny, nx = self.cube.shape
        mask = np.asarray(mask, dtype=bool)
        if mask.shape != (nchan, ny, nx):
            raise ValueError("mask must be [nchan, ny, nx]")
        if ref.xarr.v0_dv_flip[2]:
            mask = mask[::-1]                                 # the device rows run along ascending velocity
        packed = np.packbits(mask, axis=0, bitorder="little")              # [ceil(nchan / 8), ny, nx] uint8
        words = np.zeros((ny * nx, 256), dtype=np.uint8)                   # 64 words of 32 bits per pixel
        words[:, :packed.shape[0]] = packed.reshape(packed.shape[0], ny * nx).T
        bits = torch.from_numpy(words.view(np.int32)).to(eng.device)
        prob = ref._problem(self.fittype, 1, ny * nx, [-1e30] * 4, [1e30] * 4)
        out = eng.aicc_global(prob, ref.rows_on_device(), self._params_dev(1, eng), self._params_dev(2, eng),
                              expand=expand, model_f32=self.cube._data.dtype == np.float32, mask_bits=bits)
        host = torch.cat([out["rss"], out["nsamp"][:, None]], dim=1).t().contiguous().cpu().numpy()
        return host[0:3].reshape(3, ny, nx), host[3].reshape(ny, nx)

    def get_rss(self, ncomp, mask=None, planemask=None, update=True, expand=20):
        """Residual sum of squares of the ``ncomp``-component model (0: y = 0) over the dilated mask
        (UltraCube.py:432-451): ``mask`` None = the union model mask (master_model_mask), else the caller's
        [nchan, ny, nx] bool cube; ``planemask`` [ny, nx]: only these pixels of the stored maps are updated."""
        compID = str(ncomp)
        if mask is None:
            self._run_aicc(expand=expand)
            if planemask is None:
                return self.rss_maps[compID]
            return self.rss_maps[compID]          # the default maps are whole-map results already
        rss3, nsamp = self._masked_stats(mask, expand)
        with np.errstate(invalid="ignore"):
            rss = rss3[int(ncomp)].copy()
            ns = np.where(np.isnan(rss), np.nan, nsamp)
        if planemask is None or compID not in self.rss_maps:
            self.rss_maps[compID], self.NSamp_maps[compID] = rss, ns
        else:
            self.rss_maps[compID][planemask] = rss[planemask]
            self.NSamp_maps[compID][planemask] = ns[planemask]
        return self.rss_maps[compID]

    def get_AICc(self, ncomp, update=False, planemask=None, expand=20, mask=None, **kwargs):
        """UltraCube.py:474-486 (p = 4 * ncomp; ncomp = 0: model == 0, no free parameter).  With ``mask`` (the
        reference forwards it to get_rss through **kwargs) the statistics are taken over the caller's mask and stored,
        inside ``planemask`` when given, like the reference does."""
        compID = str(ncomp)
        if mask is None:
            if update and planemask is None:
                self.lnk_maps.clear()
            self._run_aicc(expand=expand)
            return self.AICc_maps[compID]
        self._run_aicc(expand=expand)              # the other maps exist before one of them is overwritten
        self.get_rss(ncomp, mask=mask, planemask=planemask, expand=expand)
        p = 4 * int(ncomp)
        with np.errstate(all="ignore"):
            rss, ns = self.rss_maps[compID], self.NSamp_maps[compID]
            fresh = ns * np.log(rss / ns) + 2 * p + (2 * p * (p + 1)) / (ns - p - 1)
        if planemask is None:
            self.AICc_maps[compID] = fresh
        else:
            self.AICc_maps[compID][planemask] = fresh[planemask]
        return self.AICc_maps[compID]

    def get_AICc_likelihood(self, ncomp1, ncomp2, **kwargs):
        """ln(L_A / L_B) = (AICc_B - AICc_A) / 2  (calc_AICc_likelihood, UltraCube.py:1120-1216)."""
        if kwargs.get("ucube_B") is not None:
            return calc_AICc_likelihood(self, ncomp1, ncomp2, ucube_B=kwargs["ucube_B"],
                                        expand=kwargs.get("expand", 0), planemask=kwargs.get("planemask"))
        self._run_aicc()
        lnk = -1.0 * (self.AICc_maps[str(ncomp1)] - self.AICc_maps[str(ncomp2)]) / 2.0
        if kwargs.get("planemask") is not None:
            # the reference updates its cached maps inside the mask only and returns those values (:1207-1214);
            # here the maps are rebuilt whole by one AICc pass whenever a fit changed, so this is just a selection
            return lnk[kwargs["planemask"]]
        return lnk

    def get_all_lnk_maps(self, ncomp_max=2, rest_model_mask=True, multicore=True):
        """UltraCube.py:1218-1280."""
        if rest_model_mask:
            self.reset_model_mask(ncomps=[2, 1], multicore=multicore)
        self._run_aicc()
        lnk10 = self.get_AICc_likelihood(1, 0)
        if ncomp_max <= 1:
            return lnk10
        return lnk10, self.get_AICc_likelihood(2, 0), self.get_AICc_likelihood(2, 1)

    def get_best_2c_parcube(self, multicore=True, lnk21_thres=5, lnk20_thres=5, lnk10_thres=5, return_lnks=True,
                            include_1c=True):
        """UltraCube.py:1283-1379.  The integer 0/1/2 selection map is kept in ``self.ncomp_map``."""
        self.reset_model_mask(ncomps=[2, 1], multicore=multicore)
        self._run_aicc(thr=(lnk10_thres, lnk20_thres, lnk21_thres))
        lnk10, lnk20, lnk21 = (self.get_AICc_likelihood(1, 0), self.get_AICc_likelihood(2, 0),
                               self.get_AICc_likelihood(2, 1))
        # same selection as the reference's masked assignments (UltraCube.py:1350-1367), as in-place masked copies
        # (np.copyto(where=...): one pass per statement, no temporaries)
        not_two = (self.ncomp_map != 2)[None]
        none = (self.ncomp_map < 1)[None]
        out = []
        for c2, c1 in ((self.pcubes['2'].parcube, self.pcubes['1'].parcube),
                       (self.pcubes['2'].errcube, self.pcubes['1'].errcube)):
            c = c2.copy()
            if include_1c:
                np.copyto(c[:4], c1[:4], where=not_two)
                np.copyto(c[4:], np.nan, where=not_two)
            else:
                np.copyto(c, np.nan, where=not_two)
            np.copyto(c, np.nan, where=none)
            out.append(c)
        parcube, errcube = out
        if return_lnks:
            lnk10[~self.has_fit(ncomp=1)] = np.nan
            m2 = self.has_fit(ncomp=2)
            lnk20[~m2] = np.nan
            lnk21[~m2] = np.nan
            return parcube, errcube, lnk10, lnk20, lnk21
        return parcube, errcube

    # ---- residual statistics ---------------------------------------------------------------------------------
    def get_residual(self, ncomp, multicore=None):
        compID = str(ncomp)
        model = self.pcubes[compID].get_modelcube()
        self.residual_cubes[compID] = get_residual(self.cube, model)
        gc.collect()
        return self.residual_cubes[compID]

    def get_rms(self, ncomp):
        self._run_resid_stats(ncomp)
        return self.rms_maps[str(ncomp)]

    def get_reduced_chisq(self, ncomp):
        self._run_resid_stats(ncomp)
        return self.rchisq_maps[str(ncomp)]

    def _run_resid_stats(self, ncomp, expand=20):
        compID = str(ncomp)
        if compID in self.rms_maps and compID in self.rchisq_maps:
            return
        p = self.pcubes[compID]
        eng = p.engine
        ny, nx = self.cube.shape[1:]
        rows = p.rows_on_device()
        prob = p._problem(self.fittype, ncomp, ny * nx, [-1e30] * (4 * ncomp), [1e30] * (4 * ncomp))
        par = self._params_dev(ncomp, eng)
        out = eng.resid_stats(prob, rows, par, expand=expand, model_f32=self.cube._data.dtype == np.float32)
        self.rms_maps[compID] = out["rms"].cpu().numpy().reshape(ny, nx)
        self.rchisq_maps[compID] = out["chisq_red"].cpu().numpy().reshape(ny, nx)


# ======================================================================================================================
class UCubePlus(UltraCube):
    """UltraCube with the file bookkeeping of a run (UltraCube.py:642-773): parameter maps live in
    ``{paraDir}/{paraNameRoot}_{n}vcomp.fits``."""

    def __init__(self, cubefile, cube=None, fittype=None, paraNameRoot=None, paraDir=None, **kwargs):
        super().__init__(cubefile, cube, fittype=fittype, **kwargs)
        self.cubeDir = os.path.dirname(cubefile)
        if paraNameRoot is None:
            self.paraNameRoot = "{}_paramaps".format(os.path.splitext(os.path.basename(cubefile))[0])
        else:
            self.paraNameRoot = paraNameRoot
        self.paraDir = "{}/para_maps".format(self.cubeDir) if paraDir is None else paraDir
        if not os.path.exists(self.paraDir):
            os.makedirs(self.paraDir)
        self.paraPaths = {}

    def _para_path(self, nc):
        if str(nc) not in self.paraPaths:
            self.paraPaths[str(nc)] = '{}/{}_{}vcomp.fits'.format(self.paraDir, self.paraNameRoot, nc)
        return self.paraPaths[str(nc)]

    def read_model_fit(self, ncomps, read_conv=False, **kwargs):
        for nc in ncomps:
            super().load_model_fit(self._para_path(nc), ncomp=nc, calc_model=False)

    def get_model_fit(self, ncomp, update=True, **kwargs):
        """Fit (``update``) and save, then load the saved maps back -- the reference's resume contract."""
        if not isinstance(ncomp, Iterable):
            ncomp = [ncomp]
        for nc in ncomp:
            self._para_path(nc)
        if update:
            for nc in ncomp:
                if 'multicore' not in kwargs:
                    kwargs['multicore'] = self.n_cores
                self.fit_cube(ncomp=[nc], **kwargs)
                self.save_fit(self.paraPaths[str(nc)], nc)
        for nc in ncomp:
            self.load_model_fit(self.paraPaths[str(nc)], nc)


def fit_cube(cube, pcube, fittype, simpfit=False, **kwargs):
    """Module-level dispatcher (UltraCube.py:777-801)."""
    if simpfit:
        return mvf.cubefit_simp(cube, pcube, fittype=fittype, **kwargs)
    return mvf.cubefit_gen(cube, pcube, fittype=fittype, **kwargs)


def calc_AICc_likelihood(ucube, ncomp_A, ncomp_B, ucube_B=None, multicore=True, expand=0, planemask=None):
    """UltraCube.py:1120-1216.  With ``ucube_B``: model A of ``ucube`` against model B of ``ucube_B`` (both on the
    same data) over their COMMON mask (model_A > 0) | (model_B > 0), dilated by ``expand`` (default 0), AICc values
    computed fresh and not stored -- the comparison replace_bad_pix makes between a new and an old fit
    (master_fitter.py:1449).  One b200fit_aicc_pair pass."""
    if ucube_B is None:
        lnk = ucube.get_AICc_likelihood(ncomp_A, ncomp_B)
        return lnk if planemask is None else lnk[planemask]
    if ncomp_A not in (0, 1, 2) or ncomp_B not in (0, 1, 2):
        raise ValueError("models of 0, 1 or 2 components can be compared")
    ref = ucube._ref_pcube if ucube._ref_pcube is not None else ucube.load_pcube()
    eng = ref.engine
    ny, nx = ucube.cube.shape[1:]
    rows = ref.rows_on_device()
    prob = ref._problem(ucube.fittype, 1, ny * nx, [-1e30] * 4, [1e30] * 4)
    # a 0-component model is y = 0 with no free parameter: an all-zero ("no fit") 1-component parameter block
    kA, kB = max(ncomp_A, 1), max(ncomp_B, 1)
    zeros = lambda: torch.zeros((ny * nx, 4), dtype=torch.float64, device=eng.device)
    pA = ucube._params_dev(ncomp_A, eng) if ncomp_A > 0 else zeros()
    pB = ucube_B._params_dev(ncomp_B, eng) if ncomp_B > 0 else zeros()
    out = eng.aicc_global(prob, rows, pA, pB, expand=expand, model_f32=ucube.cube._data.dtype == np.float32,
                          ncomp_pair=(kA, kB))
    packed = torch.cat([out["rss"], out["nsamp"][:, None]], dim=1).t().contiguous().cpu().numpy()
    with np.errstate(all="ignore"):
        aicc = []
        for k, nc in ((1, ncomp_A), (2, ncomp_B)):
            rss = packed[k].reshape(ny, nx)
            ns = np.where(np.isnan(rss), np.nan, packed[3].reshape(ny, nx))
            p = 4 * nc
            aicc.append(ns * np.log(rss / ns) + 2 * p + (2 * p * (p + 1)) / (ns - p - 1))
        lnk = -1.0 * (aicc[0] - aicc[1]) / 2.0
    return lnk if planemask is None else lnk[planemask]


def get_residual(cube, model):
    """data - model (UltraCube.py:1743-1800); NaN wherever the model cube has no fit."""
    data = cube._data if hasattr(cube, "_data") else np.asarray(cube)
    return data - model


def get_masked_moment(cube, model, order=0, expand=10, mask=None):
    """UltraCube.py:1568-1655: moment of the data over the channels where the model is positive (dilated by
    ``expand``); pixels whose model peak is below the median peak use the channels where any model exceeds 10 % of
    that median instead.  ``order`` 0, 1 or 2 with spectral_cube's definitions on the km/s axis (radio convention):
    m0 = sum(I) |dv|, m1 = sum(I v) / sum(I), m2 = sum(I (v - m1)^2) / sum(I) (a variance).  Returns [ny, nx]."""
    if order not in (0, 1, 2):
        raise ValueError("order must be 0, 1 or 2")
    data = cube._data
    with np.errstate(invalid="ignore"):
        mask = (model > 0) if mask is None else np.logical_and(mask, np.isfinite(model))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            peak_T = np.nanmax(model, axis=0)
            med_peak_T = np.nanmedian(peak_T)
        high = peak_T > med_peak_T
        mask_low = mask.copy()
        mask_low[:, high] = False
        specmask = np.any(model > med_peak_T * 0.1, axis=(1, 2))
    mask_low[specmask, :] = True
    mask[:, ~high] = mask_low[:, ~high]
    if expand > 0:
        from scipy.ndimage import binary_dilation
        mask = binary_dilation(mask, structure=np.ones((expand, 1, 1), dtype=bool))
    v = np.asarray(cube.spectral_axis, dtype=np.float64)
    keep = mask & np.isfinite(data)
    w = np.where(keep, data, 0.0)
    s0 = w.sum(axis=0, dtype=np.float64)
    empty = ~keep.any(axis=0)
    if order == 0:
        out = s0 * abs(v[1] - v[0])
    else:
        with np.errstate(all="ignore"):
            m1 = (w * v[:, None, None]).sum(axis=0, dtype=np.float64) / s0
            out = m1 if order == 1 else (w * (v[:, None, None] - m1) ** 2).sum(axis=0, dtype=np.float64) / s0
    out[empty] = np.nan
    return out
