"""Refit-loop primitives of mufasa/master_fitter.py on the B200 fitter: the pieces every refit
stage of ``master_2comp_fit`` is made of -- guesses from the best neighbour, a masked refit, the new-vs-old AICc
comparison and the in-place replacement of the better pixels -- and the first stage of the pipeline, the two-step
fit through the convolved cube.  The later stages of ``master_2comp_fit`` (``expand_fits``,
``refit_2comp_wide`` ...) are host control flow over these same primitives.

  get_refit_guesses   master_fitter.py:1512-1595  (+ utils/neighbours.py:17-84)
  replace_bad_pix     :1403-1464
  replace_para        :2471-2520
  replace_rss         :1467-1509
  Region              :65-126   (cube + parameter-file bookkeeping; logging / progress CSV left out)
  get_convolved_cube  :461-511, get_convolved_fits :516-552, get_fits :555-580, iter_2comp_fit :682-775,
  get_local_bad       :1632-1665, save_updated_paramaps :1700-1718
and the stages of ``master_2comp_fit`` (:583-679) that are compositions of the above:
  refit_bad_2comp     :777-844,  refit_marginal :847-922 (+ get_marginal_pix :1598-1629), refit_swap_2comp :925-978,
  expand_fits         :1101-1244, fit_surroundings :1247-1400, standard_2comp_fit :1668-1697,
  save_best_2comp_fit :1722-1931 (maps only: no CSV / 3-D scatter plot), get_best_2comp_model :2412-2468,
  get_best_2comp_snr_mod :2376-2409
  refit_2comp_wide    :981-1098 with get_2comp_wide_guesses :1974-2068, fit_best_2comp_residual_cnv :2071-2189,
  get_best_2comp_residual_cnv :2192-2248, get_best_2comp_residual_SpectralCube :2251-2336,
  get_best_2comp_residual :2339-2373 (+ moment_guess.mom_guess_wide_sep / window_moments / vmask_moments)
"""
import logging
import os
from copy import copy

import numpy as np
from scipy import ndimage as nd

from . import UltraCube as UCube
from . import guess_refine as gss_rf
from . import moment_guess as mmg
from . import multi_v_fit as mvf
from . import signals
from .exceptions import SNRMaskError, StartFitError
from .utils import fits_io
from .utils.convolve import convolve_gauss2d

logger = logging.getLogger(__name__)


def square_neighbour(r):
    """utils/neighbours.py: the (2r+1)^2 - 1 closest neighbours."""
    s = np.ones((2 * r + 1, 2 * r + 1), dtype=bool)
    s[r, r] = False
    return s


def maxref_neighbor_coords(mask, ref, fill_coord=(0, 0), structure=None):
    """utils/neighbours.py:17-60: for every True pixel of ``mask`` the coordinates of the neighbour (within
    ``structure``) with the highest ``ref`` value.  Neighbours outside the map are skipped the way the reference's
    IndexError handling skips them: only indices >= the array size raise; NEGATIVE indices wrap around (numpy), and
    NaN reference values win or lose ``max`` by position -- both quirks are kept."""
    if structure is None:
        structure = square_neighbour(1)
    cy, cx = int(np.ceil(structure.shape[0] / 2)) - 1, int(np.ceil(structure.shape[1] / 2)) - 1
    dy, dx = np.where(structure)
    dy, dx = dy - cy, dx - cx
    ny, nx = ref.shape
    ys, xs = np.where(mask)
    if ys.size == 0:
        return []
    if dy.size == 0:
        return [fill_coord] * ys.size
    # all pixels at once; python's sequential max() restated: the first usable neighbour stays the best when its
    # value is NaN (nothing compares greater), otherwise the first occurrence of the largest non-NaN value wins
    yy, xx = ys[:, None] + dy[None, :], xs[:, None] + dx[None, :]
    usable = (yy < ny) & (xx < nx) & (yy >= -ny) & (xx >= -nx)
    vals = np.asarray(ref)[np.where(usable, yy, 0), np.where(usable, xx, 0)]
    isnan = vals != vals
    first = np.argmax(usable, axis=1)
    rows = np.arange(ys.size)
    cand = usable & ~isnan
    with np.errstate(invalid="ignore"):
        top = np.max(np.where(cand, vals, -np.inf), axis=1)
        pick = np.argmax(cand & (vals == top[:, None]), axis=1)
    pick = np.where(isnan[rows, first], first, pick)
    has = usable.any(axis=1)
    by, bx = yy[rows, pick], xx[rows, pick]
    return [(by[i], bx[i]) if has[i] else fill_coord for i in range(ys.size)]


def get_refit_guesses(ucube, mask, ncomp, method='best_neighbour', refmap=None, structure=None):
    """master_fitter.py:1512-1595."""
    if structure is None:
        structure = square_neighbour(1)
    guesses = copy(ucube.pcubes[str(ncomp)].parcube)
    guesses[guesses == 0] = np.nan
    guesses[:, mask] = np.nan
    if method == 'best_neighbour':
        if refmap is None:
            raise ValueError("refmap must be provided for the best_neighbour method.")
        if not isinstance(refmap, np.ndarray):
            raise TypeError("{} is the incorrect type for refmap.".format(type(refmap)))
        coords = maxref_neighbor_coords(mask=mask, ref=refmap, fill_coord=(0, 0), structure=structure)
        if coords:
            ys, xs = zip(*coords)
            guesses[:, mask] = guesses[:, ys, xs]
        mask = np.logical_and(mask, np.all(np.isfinite(guesses), axis=0))
    elif method == 'convolved':
        for i, gmap in enumerate(guesses):
            gmap[mask] = np.nan
            guesses[i] = convolve_gauss2d(gmap, 2.5 / 2.355, boundary='extend')
    else:
        raise ValueError("unknown method {}".format(method))
    return guesses, mask


def replace_para(pcube, pcube_ref, mask, multicore=None):
    """master_fitter.py:2471-2520: adopt the parameters, errors, has_fit and model spectra of ``pcube_ref`` in the
    masked pixels."""
    pcube.parcube[:, mask] = pcube_ref.parcube[:, mask]
    pcube.errcube[:, mask] = pcube_ref.errcube[:, mask]
    pcube.has_fit[mask] = pcube_ref.has_fit[mask]
    if getattr(pcube, "_modelcube", None) is not None:      # the reference patches the cached model cube in place
        pix = np.flatnonzero(np.asarray(mask, bool).ravel())    # (here: only the adopted pixels are evaluated, not a
        if pix.size:                                            #  whole second model cube)
            nchan = pcube._modelcube.shape[0]
            pcube._modelcube.reshape(nchan, -1)[:, pix] = pcube_ref.model_of_pixels(pix)


def replace_rss(ucube, ucube_ref, ncomp, mask):
    """master_fitter.py:1467-1509.  The statistics of this port come from one fused AICc pass over the current
    parameter cubes, so adopting new parameters simply invalidates the cached maps; they are rebuilt on demand."""
    ucube._invalidate_stats()


def replace_bad_pix(ucube, mask, snr_min, guesses, ncomp=2, lnk21=None, simpfit=True, multicore=True,
                    return_good_mask=False):
    """master_fitter.py:1403-1464: refit the masked pixels from ``guesses`` and keep the new fit wherever it is the
    likelier one, lnk(new, old) > 0 over the common model mask (calc_AICc_likelihood with ucube_B, expand = 0)."""
    if np.sum(mask) >= 1:
        ucube_new = UCube.UltraCube(cube=ucube.cube, fittype=ucube.fittype, device=ucube.device,
                                    share_device_with=ucube)
        try:
            ucube_new.fit_cube(ncomp=[ncomp], simpfit=simpfit, maskmap=mask, snr_min=snr_min, guesses=guesses,
                               multicore=multicore)
        except SNRMaskError:
            logger.info("No valid pixel to refit with snr_min={}.".format(snr_min))
            return np.zeros((1, 1), dtype=bool) if return_good_mask else None
        lnk_NvsO = UCube.calc_AICc_likelihood(ucube_new, ncomp, ncomp, ucube_B=ucube)
        with np.errstate(invalid="ignore"):
            good_mask = np.logical_and(lnk_NvsO > 0, mask)
        logger.info("Replacing {} bad pixels with better fits".format(good_mask.sum()))
        replace_para(ucube.pcubes[str(ncomp)], ucube_new.pcubes[str(ncomp)], good_mask, multicore=multicore)
        replace_rss(ucube, ucube_new, ncomp=ncomp, mask=good_mask)
    else:
        logger.debug("not enough pixels to refit, no-refit is done")
        good_mask = np.zeros((1, 1), dtype=bool)
    if return_good_mask:
        return good_mask


# ---- the two-step fit through the convolved cube ---------------------------------------------------------------
class Region(object):
    """A cube and the files of its fits (master_fitter.py:65-126)."""

    def __init__(self, cubePath, paraNameRoot, paraDir=None, cnv_factor=2, fittype=None, multicore=True, device=None,
                 **kwargs):
        self.cubePath = cubePath
        self.paraNameRoot = paraNameRoot
        self.paraDir = paraDir
        self.fittype = 'nh3_multi_v' if fittype is None else fittype
        self.device = device
        self.ucube = UCube.UCubePlus(cubePath, paraNameRoot=paraNameRoot, paraDir=paraDir, cnv_factor=cnv_factor,
                                     fittype=self.fittype, n_cores=multicore, device=device)
        self.cnv_factor = cnv_factor
        self.progress_log = []

    def get_convolved_cube(self, **kwargs):
        get_convolved_cube(self, **kwargs)

    def get_convolved_fits(self, ncomp, **kwargs):
        get_convolved_fits(self, ncomp, **kwargs)

    def get_fits(self, ncomp, **kwargs):
        get_fits(self, ncomp, **kwargs)

    def log_progress(self, process_name, **kwargs):
        self.progress_log.append(dict(process=process_name, **kwargs))


def get_convolved_cube(reg, update=True, cnv_cubePath=None, edgetrim_width=5, paraNameRoot=None, paraDir=None,
                       multicore=True):
    """Make (or reuse) ``{cube}_conv{f}Xbeam.fits`` and open it as ``reg.ucube_cnv``."""
    root = "conv{0}Xbeam".format(int(np.rint(reg.cnv_factor)))
    reg.cnv_cubePath = "{0}_{1}.fits".format(os.path.splitext(reg.cubePath)[0], root) if cnv_cubePath is None \
        else cnv_cubePath
    reg.cnv_para_paths = {}
    if update or (not os.path.isfile(reg.cnv_cubePath)):
        reg.ucube.convolve_cube(factor=reg.cnv_factor, savename=reg.cnv_cubePath, edgetrim_width=edgetrim_width)
    if paraNameRoot is None:
        paraNameRoot = "{}_conv{}Xbeam".format(reg.paraNameRoot, int(np.rint(reg.cnv_factor)))
    if paraDir is None:
        paraDir = reg.paraDir
    reg.ucube_cnv = UCube.UCubePlus(cubefile=reg.cnv_cubePath, paraNameRoot=paraNameRoot, paraDir=paraDir,
                                    cnv_factor=reg.cnv_factor, fittype=reg.fittype, n_cores=multicore,
                                    device=reg.device)


def get_convolved_fits(reg, ncomp, update=True, **kwargs):
    kwargs = {**dict(multicore=True), **kwargs}
    if not hasattr(reg, 'ucube_cnv'):
        reg.get_convolved_cube(update=True, multicore=kwargs['multicore'])
    else:
        reg.get_convolved_cube(update=update, multicore=kwargs['multicore'])
    try:
        reg.ucube_cnv.get_model_fit(ncomp, update=update, **kwargs)
    except StartFitError as e:
        logger.warning("Fits to convolved cube failed to start. %s", e)


def get_fits(reg, ncomp, **kwargs):
    reg.ucube.get_model_fit(ncomp, **kwargs)


def get_local_bad(lnkmap, lnk_thresh=5, block_size=15, offset=0, bad_size_max=225):
    """Pixels that pass the global lnk threshold but fall in a small hole of the locally thresholded map
    (skimage threshold_local, method 'gaussian': sigma = (block_size - 1) / 6, reflected borders)."""
    local_thresh = nd.gaussian_filter(np.asarray(lnkmap, dtype=np.float64), (block_size - 1) / 6.0,
                                      mode='reflect') - offset
    with np.errstate(invalid="ignore"):
        good = (lnkmap > local_thresh) & (lnkmap > lnk_thresh)
    mask = signals.remove_small_holes(good, bad_size_max)
    return np.logical_xor(mask, good)


def save_updated_paramaps(ucube, ncomps):
    for nc in ncomps:
        if str(nc) not in ucube.paraPaths:
            ucube.paraPaths[str(nc)] = '{}/{}_{}vcomp.fits'.format(ucube.paraDir, ucube.paraNameRoot, nc)
        ucube.save_fit(ucube.paraPaths[str(nc)], nc)


def iter_2comp_fit(reg, snr_min=3.0, updateCnvFits=True, planemask=None, multicore=True, use_cnv_lnk=True,
                   save_para=True, lnk_cnv_thres=10):
    """Fit the convolved cube from moment guesses (1 and 2 components), then the cube at its own resolution with
    guesses made from the convolved fits."""
    ncomp = [1, 2]
    reg.log_progress(process_name='iter conv fits', mark_start=True)
    reg.get_convolved_fits(ncomp, update=updateCnvFits, snr_min=snr_min, multicore=multicore)
    for nc in ncomp:
        pcube_cnv = reg.ucube_cnv.pcubes[str(nc)]
        para_cnv = np.append(pcube_cnv.parcube, pcube_cnv.errcube, axis=0)
        mask = signals.remove_small_holes(np.any(np.isfinite(para_cnv), axis=0))
        para_cnv[para_cnv == 0] = np.nan
        if use_cnv_lnk:
            lnk = reg.ucube_cnv.get_AICc_likelihood(2, 1) if nc == 2 else reg.ucube_cnv.get_AICc_likelihood(1, 0)
            bad = get_local_bad(lnkmap=lnk, lnk_thresh=lnk_cnv_thres, block_size=15, offset=0, bad_size_max=225)
            with np.errstate(invalid="ignore"):
                bad = np.logical_or(bad, ~(lnk > lnk_cnv_thres))
            para_cnv[:, bad] = np.nan
            clean_map = False
        else:
            clean_map = True
        guesses = gss_rf.guess_from_cnvpara(para_cnv, reg.ucube_cnv.cube.header, reg.ucube.cube.header,
                                            clean_map=clean_map, tau_thresh=1, mask=mask)
        kwargs = {'update': True, 'guesses': guesses, 'snr_min': snr_min, 'multicore': multicore}
        n_pix = np.all(np.isfinite(guesses), axis=0).sum()
        if planemask is not None:
            kwargs['maskmap'] = planemask
            n_pix = np.sum(np.all(np.isfinite(guesses), axis=0) & planemask)
        reg.ucube.get_model_fit([nc], **kwargs)
        if save_para:
            save_updated_paramaps(reg.ucube, ncomps=[nc])
        reg.log_progress(process_name='iter %d comp fit' % nc, n_attempted=int(n_pix),
                         n_success=int(reg.ucube.pcubes[str(nc)].has_fit.sum()), finished=True)


# ---- the later stages of master_2comp_fit ----------------------------------------------------------------------
def disk_neighbour(r):
    """utils/neighbours.py: the disk(r) footprint without its centre."""
    struct = signals.disk(r)
    struct[r, r] = False
    return struct


def get_marginal_pix(lnkmap, lnk_thresh=5, holes_only=False, smallest_struct_size=9):
    """Pixels on the rim of the structures with lnk > lnk_thresh, or the holes inside them."""
    with np.errstate(invalid="ignore"):
        mask = signals.remove_small_objects(lnkmap > lnk_thresh, smallest_struct_size)
    mask_nosml = signals.remove_small_holes(mask)
    if not holes_only:
        mask_nosml = signals.binary_dilation(mask_nosml)
    return np.logical_xor(mask_nosml, mask)


def refit_bad_2comp(reg, snr_min=3, lnk_thresh=-5, multicore=True, save_para=True, method='best_neighbour',
                    local_bad=True):
    """Refit the pixels whose 2-component fit is worse than the 1-component one (or than noise) although a
    1-component fit is significant, from the parameters of their best neighbour."""
    ucube = reg.ucube
    lnk10, lnk20, lnk21 = ucube.get_all_lnk_maps(ncomp_max=2, rest_model_mask=False, multicore=multicore)
    with np.errstate(invalid="ignore"):
        mask = np.logical_or(lnk21 < lnk_thresh, lnk20 < 5)
        mask = np.logical_and(mask, lnk10 > 5)
    mask = np.logical_and(mask, np.isfinite(lnk10))
    if local_bad:
        bad_thresh = max(10, lnk_thresh)
        mask = np.logical_or(mask, get_local_bad(lnkmap=lnk21, lnk_thresh=bad_thresh, block_size=15, offset=0,
                                                 bad_size_max=225))
    if np.sum(mask) == 0:
        reg.log_progress(process_name='refit_bad_2comp', n_attempted=0, n_success=0, finished=True)
        return
    guesses, mask = get_refit_guesses(ucube, mask, ncomp=2, method='best_neighbour', refmap=lnk20)
    good_mask = replace_bad_pix(ucube, mask, snr_min, guesses, ncomp=2, simpfit=True, multicore=multicore,
                                return_good_mask=True)
    if save_para:
        save_updated_paramaps(ucube, ncomps=[2, 1])
    reg.log_progress(process_name='refit_bad_2comp', n_attempted=int(np.sum(mask)), n_success=int(good_mask.sum()),
                     finished=True)


def refit_marginal(reg, ncomp, lnk_thresh=5, holes_only=False, multicore=True, save_para=True, method='best_neighbour',
                   refit_only=True, **kwargs_marg):
    """Refit the pixels on the rim of (or in holes of) the significantly detected structures."""
    ucube = reg.ucube
    lnk_maps = ucube.get_all_lnk_maps(ncomp_max=ncomp, rest_model_mask=False, multicore=multicore)
    lnkmap = lnk_maps if ncomp == 1 else lnk_maps[2]
    proc = 'refit_marginal_%d_comp' % ncomp
    mask = get_marginal_pix(lnkmap, lnk_thresh=lnk_thresh, holes_only=holes_only, **kwargs_marg)
    if mask.sum() < 1:
        reg.log_progress(process_name=proc, n_attempted=0, n_success=0, finished=True)
        return
    guesses, mask = get_refit_guesses(ucube, mask=mask, ncomp=ncomp, method=method, refmap=lnkmap)
    if refit_only:
        mask = np.logical_and(mask, np.isfinite(guesses).all(axis=0))
        mask = np.logical_and(mask, ucube.pcubes[str(ncomp)].has_fit)
    if np.sum(mask) == 0:
        reg.log_progress(process_name=proc, n_attempted=0, n_success=0, finished=True)
        return
    good_mask = replace_bad_pix(ucube, mask, 0, guesses, ncomp=ncomp, simpfit=True, multicore=multicore,
                                return_good_mask=True)
    if save_para:
        save_updated_paramaps(ucube, ncomps=[2, 1])
    reg.log_progress(process_name=proc, n_attempted=int(np.sum(mask)), n_success=int(good_mask.sum()), finished=True)


def refit_swap_2comp(reg, snr_min=3):
    """Refit the significant 2-component pixels with the two slabs exchanged and keep the likelier fits."""
    ucube = reg.ucube
    for nc in (1, 2):
        if str(nc) not in ucube.pcubes:
            ucube.load_model_fit(ucube.paraPaths[str(nc)], nc)
    with np.errstate(invalid="ignore"):
        mask = ucube.get_AICc_likelihood(2, 1) > 5
    guesses = copy(ucube.pcubes['2'].parcube)
    guesses[0:4], guesses[4:8] = ucube.pcubes['2'].parcube[4:8].copy(), ucube.pcubes['2'].parcube[0:4].copy()
    ucube_new = UCube.UltraCube(cube=ucube.cube, fittype=reg.fittype, device=ucube.device, share_device_with=ucube)
    ucube_new.fit_cube(ncomp=[2], maskmap=mask, snr_min=snr_min, guesses=guesses)
    lnk_NvsO = UCube.calc_AICc_likelihood(ucube_new, 2, 2, ucube_B=ucube)
    with np.errstate(invalid="ignore"):
        good_mask = lnk_NvsO > 0
    replace_para(ucube.pcubes['2'], ucube_new.pcubes['2'], good_mask)
    replace_rss(ucube, ucube_new, ncomp=2, mask=good_mask)
    save_updated_paramaps(ucube, ncomps=[2, 1])


def expand_fits(reg, ncomp, lnk_thresh=5, max_iter=5, r_expand=1, fill_mask=None, lnkmap=None, multicore=True,
                snr_min=0, save_para=True, return_niter=False):
    """Grow the fitted footprint outwards from the pixels with lnk > lnk_thresh, one ring of radius ``r_expand`` per
    iteration, fitting each ring from its best neighbours; a pixel that failed twice is not tried again."""
    ucube = reg.ucube
    proc = 'expand_fit_%dcomp' % ncomp
    if lnkmap is None:
        lnkmap = ucube.get_AICc_likelihood(ncomp, ncomp - 1)
    lnkmap = np.array(lnkmap, dtype=np.float64)
    if fill_mask is None:
        fill_mask = np.ones(lnkmap.shape, dtype=bool)

    def ring(lnkmap):
        with np.errstate(invalid="ignore"):
            mask = lnkmap > lnk_thresh
        return np.logical_xor(mask, signals.binary_dilation(mask, signals.disk(r_expand))) & fill_mask

    def fit(mask):
        guesses, mask = get_refit_guesses(ucube, mask, ncomp=ncomp, method='best_neighbour', refmap=lnkmap,
                                          structure=disk_neighbour(r_expand + 1))
        good = replace_bad_pix(ucube, mask, snr_min, guesses, ncomp=ncomp, simpfit=True, multicore=multicore,
                               return_good_mask=True)
        return good if good.shape == lnkmap.shape else np.zeros(lnkmap.shape, dtype=bool)

    mask_fit = ring(lnkmap)
    mask_attempted = mask_fit
    if mask_fit.sum() < 1:
        reg.log_progress(process_name=proc, n_attempted=0, n_success=0, finished=True)
        return (0, 0, 0) if return_niter else (0, 0)
    good_mask = fit(mask_fit)
    bad_mask = np.logical_and(mask_fit, ~good_mask)
    bad_twice = np.zeros(bad_mask.shape, dtype=bool)
    n_good = n_good_total = int(good_mask.sum())
    i = 0
    while i < max_iter and n_good > 0:
        i += 1
        lnkmap[mask_fit] = ucube.get_AICc_likelihood(ncomp, ncomp - 1, planemask=mask_fit)
        mask_fit = np.logical_and(ring(lnkmap), ~mask_attempted)
        mask_fit = np.logical_or(mask_fit, bad_mask & ~bad_twice)
        if mask_fit.sum() == 0:
            break
        good_mask = fit(mask_fit)
        bad_mask = np.logical_and(mask_fit, ~good_mask)
        bad_twice = np.logical_or(bad_twice, bad_mask)
        mask_attempted = np.logical_or(mask_attempted, mask_fit)
        n_good = int(good_mask.sum())
        n_good_total += n_good
    if save_para:
        save_updated_paramaps(ucube, ncomps=[2, 1])
    n_attempt_total = int(mask_attempted.sum())
    reg.log_progress(process_name=proc, n_attempted=n_attempt_total, n_success=n_good_total, finished=True)
    return (n_good_total, n_attempt_total, i) if return_niter else (n_good_total, n_attempt_total)


def fit_surroundings(reg, ncomps=[1, 2], snr_min=3, max_iter=None, save_para=True, fill_mask=None, multicore=True,
                     match_footprint=True):
    """``expand_fits`` in tiers of falling lnk thresholds (10, 3, then 0 / -5 / -20 with wider rings as ``snr_min``
    allows), the lowest ncomp first; with ``match_footprint`` a higher ncomp may only grow where the previous one
    has fits.  Small holes left in the footprint get one more pass."""
    max_iter_fill = 100
    if max_iter is None:
        max_iter = 1 if snr_min > 3 else max_iter_fill
    elif max_iter > max_iter_fill:
        max_iter_fill = max_iter
    if fill_mask is None:
        fill_mask = reg.ucube.pcubes[str(min(ncomps))].footprint() if str(min(ncomps)) in reg.ucube.pcubes \
            else np.any(reg.ucube.cube.mask.include(), axis=0)
    ncomps = sorted(ncomps)
    for i, ncomp in enumerate(ncomps):
        m_iter = max_iter
        if match_footprint and i > 0:
            m_iter = max_iter_fill
            prev = reg.ucube.has_fit(ncomps[i - 1])
            fill_mask = np.logical_or(fill_mask, prev) if i > 1 else prev
        kw = dict(ncomp=ncomp, save_para=save_para, fill_mask=fill_mask, multicore=multicore, return_niter=True,
                  snr_min=snr_min)
        state = dict(m_iter=m_iter, n_good=1, good=0, tried=0)

        def tier(lnk_thresh, r_expand):
            if state["m_iter"] > 0 and state["n_good"] > 0:
                n_good, n_attempt, n = expand_fits(reg, lnk_thresh=lnk_thresh, max_iter=state["m_iter"],
                                                   r_expand=r_expand, **kw)
                state["n_good"] = n_good
                state["good"] += n_good
                state["tried"] += n_attempt
                state["m_iter"] -= n

        tier(10, 1)
        tier(3, 1)
        last = (3, 1)
        for cut, thr, r in ((2, 0, 2), (1, -5, 3), (0, -20, 5)):
            if snr_min <= cut:
                tier(thr, r)
                last = (thr, r)
        fitted = reg.ucube.has_fit(ncomp)
        filled = signals.remove_small_holes(fitted, 9)
        if np.logical_xor(fitted, filled).sum() > 0:
            kw['fill_mask'] = filled
            state["n_good"], state["m_iter"] = 1, 1
            tier(last[0], 2)
        reg.log_progress(process_name='expand_fit_%dcomp' % ncomp, n_attempted=state["tried"],
                         n_success=state["good"], finished=True)


def standard_2comp_fit(reg, planemask=None, snr_min=3):
    """1- and 2-component fits from moment guesses only (no convolved-cube stage)."""
    for nc in (1, 2):
        kwargs = {'update': True, 'snr_min': snr_min}
        if planemask is not None:
            kwargs['maskmap'] = planemask
        reg.ucube.get_model_fit([nc], **kwargs)


def get_best_2comp_model(reg):
    """[nchan, ny, nx] model cube of the selected fit per pixel: 2 components where lnk21 > 5 and lnk20 > 5, else 1
    where lnk10 > 5, else 0."""
    lnk10, lnk20, lnk21 = reg.ucube.get_all_lnk_maps(ncomp_max=2, rest_model_mask=True)
    mod1 = reg.ucube.pcubes['1'].get_modelcube()
    mod2 = reg.ucube.pcubes['2'].get_modelcube()
    modbest = np.zeros_like(mod1)
    with np.errstate(invalid="ignore"):
        mask = lnk10 > 5
        modbest[:, mask] = mod1[:, mask]
        mask = np.logical_and(lnk21 > 5, lnk20 > 5)
    modbest[:, mask] = mod2[:, mask]
    return modbest


def _rms_of_residual(res):
    """UltraCube.get_rms (UltraCube.py:1701-1740): 1.4826 median|r - roll(r, 2)| / sqrt(2) along the spectrum."""
    diff = res - np.roll(res, 2, axis=0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return 1.4826 * np.nanmedian(np.abs(diff), axis=0) / 2 ** 0.5


def get_best_2comp_snr_mod(reg, modbest=None):
    """Peak of the selected model over the noise of its residual."""
    if modbest is None:
        modbest = get_best_2comp_model(reg)
    rms = _rms_of_residual(reg.ucube.cube._data - modbest)
    import warnings
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmax(modbest, axis=0) / rms


def save_map(map, header, savename, overwrite=True):
    if not overwrite and os.path.exists(savename):
        raise IOError("%s exists" % savename)
    fits_io.write(savename, np.asarray(map), header)


def _best_model_products(reg, expand=20):
    """SNR, model moment 0 and masked data moment 0 of the selected model, from the device-resident rows.

    The reference builds three [nchan, ny, nx] host cubes for these maps (two model cubes + the selected one,
    master_fitter.py:2412-2468, then get_masked_moment, UltraCube.py:1568-1655); at 512 x 512 x 1000 that is minutes
    of numpy.  Here the 1- and 2-component models are evaluated by ``b200fit_model`` into pixel-major rows, the
    per-pixel choice is the integer map of the AICc kernel, and what is left are row reductions on the device (torch
    tensor ops: plumbing around the kernels, off the fit path)."""
    import torch
    ucube = reg.ucube
    ucube.get_all_lnk_maps(ncomp_max=2, rest_model_mask=True)
    ref = ucube._ref_pcube if ucube._ref_pcube is not None else ucube.load_pcube()
    eng = ref.engine
    nchan, ny, nx = ucube.cube.shape
    npix = ny * nx
    rows = ref.rows_on_device()[:, :nchan]
    ncomp = torch.from_numpy(ucube.ncomp_map.astype(np.int64).ravel()).to(eng.device)
    best = torch.zeros((npix, nchan), dtype=torch.float64, device=eng.device)
    for nc in (1, 2):
        sel = torch.nonzero(ncomp == nc).ravel()
        if sel.numel() == 0:
            continue
        par = ucube._params_dev(nc, eng)[sel].contiguous()
        prob = ref._problem(ucube.fittype, nc, int(sel.numel()), [-1e30] * (4 * nc), [1e30] * (4 * nc))
        best[sel] = eng.model(prob, par)[:, :nchan]
    best = torch.nan_to_num(best, nan=0.0)
    peak = best.amax(dim=1)
    v = np.asarray(ucube.cube.spectral_axis, dtype=np.float64)
    dv = abs(v[1] - v[0])
    # S/N: peak of the selected model over the noise of its residual (get_rms of the selected fit's residual)
    rms = torch.full((npix,), float("nan"), dtype=torch.float64, device=eng.device)
    for nc in (1, 2):
        m = torch.from_numpy(np.ascontiguousarray(ucube.get_rms(nc)).ravel()).to(eng.device)
        rms = torch.where(ncomp == nc, m, rms)
    snr = torch.where(ncomp > 0, peak / rms, torch.zeros_like(peak))
    model_mom0 = best.sum(dim=1) * dv
    # get_masked_moment: channels where the model is positive; pixels fainter than the median peak use the channels
    # where any model exceeds 10 % of that median; dilated by ones(expand) -> [k - expand/2, k + expand/2 - 1]
    mask = best > 0
    med = torch.nanmedian(peak)
    high = peak > med
    spec = (best > 0.1 * med).any(dim=0)
    mask = torch.where(high[:, None], mask, mask | spec[None, :])
    if expand > 0:
        lo, hi = expand // 2, expand - expand // 2 - 1
        padded = torch.nn.functional.pad(mask.to(torch.float32)[:, None, :], (hi, lo))
        mask = torch.nn.functional.max_pool1d(padded, kernel_size=expand, stride=1)[:, 0, :] > 0
    keep = mask & torch.isfinite(rows)
    mom0 = torch.where(keep, rows.to(torch.float64), torch.zeros((), dtype=torch.float64, device=eng.device)).sum(1) * dv
    mom0 = torch.where(keep.any(dim=1), mom0, torch.full_like(mom0, float("nan")))
    _, _, flip = ref.xarr.v0_dv_flip           # rows are stored with ascending velocity; sums do not care
    out = torch.stack([snr, model_mom0, mom0]).cpu().numpy().reshape(3, ny, nx)
    return dict(SNR=out[0], model_mom0=out[1], mom0=out[2])


def save_best_2comp_fit(reg, multicore=True, from_saved_para=False, lnk21_thres=5, lnk10_thres=5, save_csv=False,
                        save_Scatter3D=False):
    """The products of a run: ``{root}_2vcomp_final.fits`` (model-selected parameters + errors), the lnk21 / lnk10 /
    lnk20 maps, the peak S/N of the selected model, moment 0 of the model and of the data over the model's
    footprint, reduced chi-squared of the 1- and 2-component fits.  (The reference's CSV table and 3-D scatter plot
    are plotting-side products and are not written.)"""
    if save_csv or save_Scatter3D:
        raise NotImplementedError("CSV / 3-D scatter products are outside the fit path")
    ncomps = [1, 2]
    ucube = reg.ucube
    save_updated_paramaps(ucube, ncomps=ncomps)
    if from_saved_para:
        for nc in ncomps:
            ucube.load_model_fit(ucube.paraPaths[str(nc)], nc)
    pcube_final = ucube.pcubes['2'].copy()
    parcube, errcube, lnk10, lnk20, lnk21 = ucube.get_best_2c_parcube(multicore=multicore, lnk21_thres=lnk21_thres,
                                                                     lnk10_thres=lnk10_thres, return_lnks=True)
    pcube_final.parcube, pcube_final.errcube = parcube, errcube
    root2 = os.path.splitext(ucube._para_path(2))[0]
    mvf.save_pcube(pcube_final, "{}_final.fits".format(root2), ncomp=2,
                   header_note='Model-selected best 1- or 2-comp fits parameters, based on lnk21')
    paraDir, paraRoot = ucube.paraDir, ucube.paraNameRoot

    def name(key):
        for word in ("parameters", "parameter", "para"):
            if word in paraRoot:
                return "{}/{}.fits".format(paraDir, paraRoot.replace(word, key))
        return "{}/{}_{}.fits".format(paraDir, paraRoot, key)

    def header(notes, bunit=None):
        hdr = mvf.make_header(2, ucube.cube.header)
        hdr.set('NOTES', notes)
        if bunit is not None:
            hdr.set('BUNIT', bunit)
        return hdr

    for key, lnk in (("21", lnk21), ("10", lnk10), ("20", lnk20)):
        other = 'noise' if key[1] == '0' else '%s comp fit' % key[1]
        save_map(lnk, header("Relative log-likelihood of %s comp fit vs %s" % (key[0], other), 'unitless'),
                 name("lnk" + key))
    maps = _best_model_products(reg, expand=20)
    save_map(maps["SNR"], header('Estimated peak signal-to-noise ratio', 'unitless'), name("SNR"))
    save_map(maps["model_mom0"], header('moment 0 of the best-fit model', 'K km/s'), name("model_mom0"))
    save_map(maps["mom0"], header('moment 0 over the footprint of the best-fit model', 'K km/s'), name("mom0"))
    save_map(ucube.get_reduced_chisq(1), header('reduced chi-squared values of the 1-comp model'), name("chi2red_1c"))
    save_map(ucube.get_reduced_chisq(2), header('reduced chi-squared values of the 2-comp model'), name("chi2red_2c"))
    return reg


# ---- recovery of a second component at wide velocity separation ----------------------------------------------
# The reference masks the residual cube in get_best_2comp_residual_SpectralCube with the branches the wrong way
# round (master_fitter.py:2322-2332): where at least one pixel of the residual exceeds ``res_snr_cut`` the cube is
# returned WITHOUT a mask, and get_best_2comp_residual_cnv then raises SNRMaskError because ``cube.mask is None``
# (:2238-2242) -- refit_2comp_wide logs "No second component recovered" and returns; only a residual with NO pixel
# above the cut gets a (trivial, all-inclusive) mask and goes on to the convolved fit.  A drop-in has to behave the
# same, so that is the default here.  RESIDUAL_MASK_AS_DOCUMENTED = True instead does what the reference's docstrings
# describe: the cube is masked to the (dilated) pixels whose residual peak exceeds the cut, and left unmasked -- but
# still fitted -- when there are none.
RESIDUAL_MASK_AS_DOCUMENTED = False


def get_skyheader(cube_header):
    """master_fitter.py:2523-2559: the celestial cards of a cube header."""
    from .convolve_tools import get_celestial_hdr
    hdr2D = get_celestial_hdr(cube_header)
    hdr2D['NAXIS'] = 2
    return hdr2D


def _best_residual_rows(reg):
    """(residual [npix, nchan] float32 on the device, velocity ascending; engine; flip): data minus the selected model
    per pixel, from the device-resident rows and ``b200fit_model`` -- the reference builds two host model cubes, the
    selected one and their difference (master_fitter.py:2339-2373, 2412-2468; half a minute of numpy at
    512 x 512 x 1000, and a cached model cube that every later ``replace_para`` would have to patch)."""
    import torch
    ucube = reg.ucube
    ucube.get_all_lnk_maps(ncomp_max=2, rest_model_mask=True)
    ref = ucube._ref_pcube if ucube._ref_pcube is not None else ucube.load_pcube()
    eng = ref.engine
    nchan, ny, nx = ucube.cube.shape
    res = ref.rows_on_device()[:, :nchan].clone()
    ncomp = torch.from_numpy(ucube.ncomp_map.astype(np.int64).ravel()).to(eng.device)
    for nc in (1, 2):
        sel = torch.nonzero(ncomp == nc).ravel()
        if sel.numel() == 0:
            continue
        par = ucube._params_dev(nc, eng)[sel].contiguous()
        prob = ref._problem(ucube.fittype, nc, int(sel.numel()), [-1e30] * (4 * nc), [1e30] * (4 * nc))
        # the host recipe subtracts the model cube in the data's precision (get_modelcube has the cube's dtype)
        model = torch.nan_to_num(eng.model(prob, par)[:, :nchan], nan=0.0).to(torch.float32)
        res[sel] -= model
    _, _, flip = ref.xarr.v0_dv_flip
    return res, eng, flip


def _rows_to_cube(rows, ny, nx, flip):
    """[npix, nchan] device rows (velocity ascending) -> [nchan, ny, nx] numpy in the data cube's channel order."""
    import torch
    if flip:
        rows = torch.flip(rows, dims=[1])
    return np.ascontiguousarray(rows.t().cpu().numpy()).reshape(rows.shape[1], ny, nx)


def _rms_of_rows(res):
    """UltraCube.get_rms on device rows: 1.4826 median|r - roll(r, 2)| / sqrt(2) over the finite differences (numpy's
    median: the mean of the two middle values of an even count)."""
    import torch
    diff = (res - torch.roll(res, 2, dims=1)).abs().to(torch.float64)
    n = torch.isfinite(diff).sum(dim=1)
    srt = torch.sort(torch.where(torch.isfinite(diff), diff, torch.full_like(diff, float("inf"))), dim=1).values
    lo = torch.clamp((n - 1) // 2, min=0)[:, None]
    hi = torch.clamp(n // 2, min=0)[:, None]
    med = 0.5 * (torch.gather(srt, 1, lo) + torch.gather(srt, 1, hi))[:, 0]
    med = torch.where(n > 0, med, torch.full_like(med, float("nan")))
    return 1.4826 * med / 2 ** 0.5


def get_best_2comp_residual(reg):
    """master_fitter.py:2339-2373: data minus the selected model, as a cube on the data's axes."""
    res, eng, flip = _best_residual_rows(reg)
    nchan, ny, nx = reg.ucube.cube.shape
    return reg.ucube.cube.with_data(_rows_to_cube(res, ny, nx, flip))


def get_best_2comp_residual_SpectralCube(reg, masked=True, window_hwidth=3.5, res_snr_cut=5):
    """master_fitter.py:2251-2336.  ``masked``: pixels are selected by the peak of the residual within
    ``window_hwidth`` of the 1-component velocity over the residual's own noise (see RESIDUAL_MASK_AS_DOCUMENTED for
    what the selection then does)."""
    import torch
    res, eng, flip = _best_residual_rows(reg)
    nchan, ny, nx = reg.ucube.cube.shape
    cube_res = reg.ucube.cube.with_data(_rows_to_cube(res, ny, nx, flip))
    if masked and res_snr_cut > 0:
        best_rms = _rms_of_rows(res)
        vmap = torch.from_numpy(np.ascontiguousarray(reg.ucube.pcubes['1'].parcube[0], dtype=np.float64).ravel()
                                ).to(eng.device)
        v = torch.from_numpy(np.sort(np.asarray(reg.ucube.cube.spectral_axis, dtype=np.float64))).to(eng.device)
        window = (v[None, :] > (vmap - window_hwidth)[:, None]) & (v[None, :] < (vmap + window_hwidth)[:, None])
        inside = window & torch.isfinite(res)
        peak = torch.where(inside, res, torch.full_like(res, -float("inf"))).amax(dim=1).to(torch.float64)
        peak = torch.where(inside.any(dim=1), peak, torch.full_like(peak, float("nan")))
        mask_res = ((peak / best_rms) > res_snr_cut).cpu().numpy().reshape(ny, nx)
        if RESIDUAL_MASK_AS_DOCUMENTED:
            if mask_res.sum() >= 1:
                return cube_res.with_mask(signals.binary_dilation(mask_res))
            logger.debug("No pixel in the residual cube mask. No mask will be applied.")
            return cube_res.with_mask(np.ones(mask_res.shape, bool))
        if mask_res.sum() < 1:
            logger.debug("No pixel in the residual cube mask. No mask will be applied.")
            return cube_res.with_mask(~signals.binary_dilation(mask_res))
    return cube_res


def get_best_2comp_residual_cnv(reg, masked=True, window_hwidth=3.5, res_snr_cut=3):
    """master_fitter.py:2192-2248: the residual cube smoothed and regridded by the region's convolution factor."""
    from . import convolve_tools as cnvtool
    cube_res_masked = get_best_2comp_residual_SpectralCube(reg, masked=masked, window_hwidth=window_hwidth,
                                                           res_snr_cut=res_snr_cut)
    if cube_res_masked.plane_include is None:            # the reference's ``cube.mask is None``
        msg = "Cube_res_masked does not have a mask for res_snr_cut={}. No residual refitting will be " \
              "performed".format(res_snr_cut)
        logger.debug(msg)
        raise SNRMaskError(msg)
    return cnvtool.convolve_sky_byfactor(cube_res_masked, factor=reg.cnv_factor, edgetrim_width=None, snrmasked=False,
                                         iterrefine=False, device=reg.device)


def fit_best_2comp_residual_cnv(reg, window_hwidth=3.5, res_snr_cut=3, savefit=True, planemask=None):
    """master_fitter.py:2071-2189: a 1-component fit of the convolved residual (``reg.ucube_res_cnv``) from windowed
    moment guesses about the (convolved, or median-smoothed and regridded) 1-component velocity; pixels enter with
    amplitude / mad_std >= ``res_snr_cut``."""
    from scipy.signal import medfilt2d
    from . import convolve_tools as cnvtool
    cube_res_cnv = get_best_2comp_residual_cnv(reg, masked=True, window_hwidth=window_hwidth, res_snr_cut=res_snr_cut)
    ncomp = 1
    if not hasattr(reg, 'ucube_cnv'):
        pcube1 = reg.ucube.pcubes['1']
        vmap = medfilt2d(np.asarray(pcube1.parcube[0], dtype=np.float64), kernel_size=3)
        vmap = cnvtool.regrid(vmap, header1=get_skyheader(reg.ucube.cube.header),
                              header2=get_skyheader(cube_res_cnv.header))
    else:
        vmap = reg.ucube_cnv.pcubes['1'].parcube[0]
    if vmap.shape != cube_res_cnv.shape[1:]:
        raise SNRMaskError("There doesn't seem to be enough pixels to find the residual moments. the velocity map "
                           "{} does not match the convolved residual {}".format(vmap.shape, cube_res_cnv.shape[1:]))
    m0, m1, m2 = mmg.window_moments(cube_res_cnv, v_atpeak=vmap, window_hwidth=window_hwidth, device=reg.device)
    gg = mmg.moment_guesses_1c(m0, m1, m2)
    gg = gss_rf.refine_each_comp(gg, mask=np.isfinite(gg[0]))
    if res_snr_cut > 0:
        rms = mmg.mad_std_cube(cube_res_cnv._data, device=reg.device)
        with np.errstate(invalid="ignore", divide="ignore"):
            maskmap = m0 / rms >= res_snr_cut
    else:
        maskmap = np.isfinite(m0)
    if planemask is not None:
        if planemask.shape == maskmap.shape:
            maskmap = np.logical_and(maskmap, planemask)
        else:
            # reproject_interp (bilinear) of the dilated mask onto the convolved grid, thresholded at 0.5
            from .utils.fits_utils import linear_pixel_map
            grown = signals.binary_dilation(planemask, signals.disk(2)).astype(np.float64)
            (ay, by), (ax, bx) = linear_pixel_map(reg.ucube.cube.header, cube_res_cnv.header)
            y = ay * np.arange(maskmap.shape[0]) + by
            x = ax * np.arange(maskmap.shape[1]) + bx
            inside = np.logical_and.outer((y >= 0) & (y <= grown.shape[0] - 1), (x >= 0) & (x <= grown.shape[1] - 1))
            y0 = np.clip(np.floor(y).astype(int), 0, grown.shape[0] - 2)
            x0 = np.clip(np.floor(x).astype(int), 0, grown.shape[1] - 2)
            fy, fx = np.clip(y - y0, 0, 1)[:, None], np.clip(x - x0, 0, 1)[None, :]
            interp = ((1 - fy) * ((1 - fx) * grown[np.ix_(y0, x0)] + fx * grown[np.ix_(y0, x0 + 1)]) +
                      fy * ((1 - fx) * grown[np.ix_(y0 + 1, x0)] + fx * grown[np.ix_(y0 + 1, x0 + 1)]))
            maskmap = np.logical_and(maskmap, inside & (interp > 0.5))
    reg.ucube_res_cnv = UCube.UltraCube(cube=cube_res_cnv, fittype=reg.fittype, device=reg.device)
    if np.sum(maskmap) > 0:
        try:
            reg.ucube_res_cnv.fit_cube(ncomp=[1], simpfit=False, snr_min=0.0, guesses=gg, maskmap=maskmap)
        except StartFitError:
            raise StartFitError("The first fitted pixel to the convovled residual cube did not yield a fit, likely due "
                                "to a lack of signal or poor guesses.")
    else:
        raise SNRMaskError("No valid pixel to refit in the residual cube with snr_min={}. Please consider trying a "
                           "lower snr_min value".format(res_snr_cut))
    if savefit:
        oriParaPath = reg.ucube.paraPaths[str(ncomp)]
        reg.ucube_res_cnv.save_fit("{}_onBestResidual.fits".format(os.path.splitext(oriParaPath)[0]), ncomp)


def get_2comp_wide_guesses(reg, window_hwidth=3.5, snr_min=3, savefit=True, planemask=None):
    """master_fitter.py:1974-2068: [4, ny, nx] guesses for the second, widely separated component -- the
    1-component fit of the convolved residual where it beats noise (lnk10 > 5), regridded to the full resolution;
    moment guesses of the unmasked residual about the 1-component velocity otherwise."""
    from scipy.signal import medfilt2d
    if not hasattr(reg, 'ucube_res_cnv'):
        try:
            fit_best_2comp_residual_cnv(reg, window_hwidth=window_hwidth, res_snr_cut=snr_min, savefit=savefit,
                                        planemask=planemask)
        except SNRMaskError:
            raise SNRMaskError("No valid pixel to refit in the residual cube for snr_min={}.".format(snr_min))

    def get_mom_guesses(reg):
        cube_res = get_best_2comp_residual_SpectralCube(reg, masked=False, window_hwidth=3.5)
        vmap = medfilt2d(np.asarray(reg.ucube.pcubes['1'].parcube[0], dtype=np.float64), kernel_size=3)
        moms_res = mmg.vmask_moments(cube_res, vmap=vmap, window_hwidth=3.5, device=reg.device)
        gg = mmg.moment_guesses_1c(moms_res[0], moms_res[1], moms_res[2])
        return gss_rf.refine_each_comp(gg, mask=np.isfinite(moms_res[0]))

    if not hasattr(reg, 'ucube_res_cnv') or ('1' not in reg.ucube_res_cnv.pcubes) or not hasattr(
            reg.ucube_res_cnv.pcubes['1'], 'parcube'):
        return get_mom_guesses(reg)
    with np.errstate(invalid="ignore"):
        aic1v0_mask = reg.ucube_res_cnv.get_AICc_likelihood(1, 0) > 5
    if np.sum(aic1v0_mask) >= 1:
        logger.info("number of good fits to convolved residual: {}".format(np.sum(aic1v0_mask)))
        pc = reg.ucube_res_cnv.pcubes['1']
        preguess = np.append(pc.parcube, pc.errcube, axis=0).copy()
        preguess[:, ~aic1v0_mask] = np.nan
        gmask = signals.binary_dilation(aic1v0_mask)
        return gss_rf.guess_from_cnvpara(preguess, reg.ucube_res_cnv.cube.header, reg.ucube.cube.header, mask=gmask)
    logger.info("no good fit from convolved guess, using the moment guess for the full-res refit instead")
    return get_mom_guesses(reg)


def refit_2comp_wide(reg, snr_min=3, method='residual', planemask=None, multicore=True, save_para=True):
    """master_fitter.py:981-1098: where the 1-component fit beat both noise and the 2-component fit (lnk10 > 5,
    lnk21 < -5), look for a second component far from the first -- ``method`` 'residual': in the convolved residual
    of the best model (get_2comp_wide_guesses), 'moments': in the moments of the spectrum with its main component
    blanked (mom_guess_wide_sep) -- refit with both and keep the refits that are likelier than the old fit."""
    proc_name = 'refit_2comp_wide'
    reg.log_progress(process_name=proc_name, mark_start=True)
    for nc in (1, 2):
        if str(nc) not in reg.ucube.pcubes:
            if str(nc) not in reg.ucube.paraPaths:
                reg.ucube.paraPaths[str(nc)] = '{}/{}_{}vcomp.fits'.format(reg.ucube.paraDir, reg.ucube.paraNameRoot, nc)
            reg.ucube.load_model_fit(reg.ucube.paraPaths[str(nc)], nc)
    lnk10, lnk20, lnk21 = reg.ucube.get_all_lnk_maps(ncomp_max=2, rest_model_mask=False)
    if planemask is None:
        with np.errstate(invalid="ignore"):
            mask = np.logical_and(lnk21 < -5, lnk10 > 5)
    else:
        mask = planemask
    if method == 'residual':
        c1_guess = gss_rf.refine_each_comp(copy(reg.ucube.pcubes['1'].parcube))
        try:
            wide_comp_guess = get_2comp_wide_guesses(reg, window_hwidth=3.5, snr_min=snr_min, savefit=True,
                                                     planemask=mask)
        except SNRMaskError as e:
            logger.warning(str(e) + " No second component recovered from the residual cube.")
            reg.log_progress(process_name=proc_name, mark_start=False, n_attempted=0, n_success=0, finished=True)
            return
        except StartFitError as e:
            logger.warning(str(e))
            return
        except ValueError as e:
            logger.warning(str(e) + "Convolved residual cube is not fitted. No recovery is done for the wide "
                                    "component.")
            reg.log_progress(process_name=proc_name, mark_start=False, n_attempted=0, n_success=0, finished=True)
            return
        wide_comp_guess[:, ~mask] = np.nan
        final_guess = np.append(c1_guess, wide_comp_guess, axis=0)
        mask = np.logical_and(mask, np.all(np.isfinite(final_guess), axis=0))
        final_guess = gss_rf.refine_2c_guess(final_guess)
    elif method == 'moments':
        final_guess = mmg.mom_guess_wide_sep(reg.ucube.cube, planemask=mask, device=reg.device)
        mask = np.logical_and(mask, np.all(np.isfinite(final_guess), axis=0))
    else:
        raise Exception("the following method specified is invalid: {}".format(method))
    mask_size = int(np.sum(mask))
    if mask_size > 0:
        logger.info("Attempting wide recovery over {} pixels".format(mask_size))
        reg.log_progress(process_name=proc_name, mark_start=False, n_attempted=mask_size, n_success='in progress')
    else:
        logger.info("No pixel was used in attempt to recover wide component")
        reg.log_progress(process_name=proc_name, mark_start=False, n_attempted=0, n_success=None, finished=True)
        return
    good_mask = replace_bad_pix(reg.ucube, mask, snr_min, final_guess, ncomp=2, simpfit=True, multicore=multicore,
                                return_good_mask=True)
    if save_para:
        save_updated_paramaps(reg.ucube, ncomps=[2, 1])
    reg.log_progress(process_name=proc_name, mark_start=False, n_attempted=None, n_success=int(good_mask.sum()),
                     finished=True)
    return good_mask


def master_2comp_fit(reg, snr_min=0.0, recover_wide=True, planemask=None, updateCnvFits=True, refit_bad_pix=True,
                     refit_marg=True, max_expand_iter=None, multicore=True):
    """The reference's full 2-component pipeline (master_fitter.py:583-679): two-step fit through the convolved cube,
    quality-assurance refits, recovery of widely separated components, expansion into the surroundings, products."""
    iter_2comp_fit(reg, snr_min=snr_min, updateCnvFits=updateCnvFits, planemask=planemask, multicore=multicore)
    recover_snr_min = min(3.0, snr_min / 3.0)
    if refit_bad_pix:
        refit_bad_2comp(reg, snr_min=recover_snr_min, lnk_thresh=-5, multicore=multicore)
    if recover_wide:
        refit_2comp_wide(reg, snr_min=recover_snr_min, multicore=multicore)
    if refit_marg:
        for nc in (1, 2):
            refit_marginal(reg, ncomp=nc, lnk_thresh=10, holes_only=False, multicore=multicore,
                           method='best_neighbour')
    if max_expand_iter is not False:
        if max_expand_iter is True:
            max_expand_iter = None
        fit_surroundings(reg, ncomps=[1, 2], snr_min=snr_min, max_iter=max_expand_iter, fill_mask=None,
                         multicore=multicore, match_footprint=True)
        if refit_bad_pix:
            refit_bad_2comp(reg, snr_min=recover_snr_min, lnk_thresh=-5, multicore=multicore)
    save_best_2comp_fit(reg, multicore=multicore, lnk21_thres=5, lnk10_thres=5)
    return reg
