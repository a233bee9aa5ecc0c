"""Refit-loop primitives of mufasa/master_fitter.py on the B200 fitter: the pieces every refit
stage of ``master_2comp_fit`` is made of -- guesses from the best neighbour, a masked refit, the new-vs-old AICc
comparison and the in-place replacement of the better pixels.  The multi-stage control flow itself (``Region``,
``iter_2comp_fit``, ``expand_fits`` ...) is host logic on 2-D maps and is not rebuilt here.

  get_refit_guesses   master_fitter.py:1512-1595  (+ utils/neighbours.py:17-84)
  replace_bad_pix     :1403-1464
  replace_para        :2471-2520
  replace_rss         :1467-1509
"""
import logging
from copy import copy

import numpy as np

from . import UltraCube as UCube
from .exceptions import SNRMaskError
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
