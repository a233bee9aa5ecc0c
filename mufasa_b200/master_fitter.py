"""Refit-loop primitives of mufasa/master_fitter.py on the B200 fitter: the pieces every refit
stage of ``master_2comp_fit`` is made of -- guesses from the best neighbour, a masked refit, the new-vs-old AICc
comparison and the in-place replacement of the better pixels -- and the first stage of the pipeline, the two-step
fit through the convolved cube.  The later stages of ``master_2comp_fit`` (``expand_fits``,
``refit_2comp_wide`` ...) are host control flow over these same primitives and are not rebuilt here.

  get_refit_guesses   master_fitter.py:1512-1595  (+ utils/neighbours.py:17-84)
  replace_bad_pix     :1403-1464
  replace_para        :2471-2520
  replace_rss         :1467-1509
  Region              :65-126   (cube + parameter-file bookkeeping; logging / progress CSV left out)
  get_convolved_cube  :461-511, get_convolved_fits :516-552, get_fits :555-580, iter_2comp_fit :682-775,
  get_local_bad       :1632-1665, save_updated_paramaps :1700-1718
"""
import logging
import os
from copy import copy

import numpy as np
from scipy import ndimage as nd

from . import UltraCube as UCube
from . import guess_refine as gss_rf
from . import signals
from .exceptions import SNRMaskError, StartFitError
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
    out = []
    for y, x in zip(*np.where(mask)):
        best, best_val = None, None
        for yy, xx in zip(dy + y, dx + x):
            if yy >= ny or xx >= nx or yy < -ny or xx < -nx:
                continue
            val = ref[yy, xx]
            if best is None or val > best_val:      # python max(): the first of equal / incomparable values stays
                best, best_val = (yy, xx), val
        out.append(best if best is not None else fill_coord)
    return out


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
        newmod = pcube_ref.get_modelcube(update=False)
        pcube._modelcube[:, mask] = newmod[:, mask]


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
