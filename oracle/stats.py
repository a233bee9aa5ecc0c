"""Statistics / model-selection half of the hot path (oracle).  TEST INFRASTRUCTURE.

Restates on plain numpy arrays (``cube`` = ndarray[nchan, ny, nx]):
  * aic.AIC / AICc / likelihood                         mufasa/aic.py:136-201
  * UltraCube.expand_mask                               mufasa/UltraCube.py:1658-1698
  * UltraCube.get_rss (include_nosamp fill, dilation)   mufasa/UltraCube.py:1384-1475
  * UltraCube.get_rms / get_chisq                       mufasa/UltraCube.py:1701-1740, 1478-1565
  * UltraCube.get_rss method (nsamp<1, rss<=0 -> NaN)   mufasa/UltraCube.py:432-451
  * UltraCube.get_AICc (p = 4*ncomp)                    mufasa/UltraCube.py:474-486
  * get_all_lnk_maps / get_best_2c_parcube / has_fit    mufasa/UltraCube.py:1218-1280, 1283-1379, 328-348
Pinned against the reference's own code by tests/golden/make_golden.py (stats_*.npz).
"""
from copy import copy

import numpy as np
from scipy import ndimage as nd


def AIC(rss, p, N):
    N = np.array(N, dtype=float)        # the reference mutates N in place; same values result
    N[N == 0] = np.nan
    return N * np.log(rss / N) + 2 * p


def AICc(rss, p, N):
    top = 2 * p * (p + 1)
    N = np.array(N, dtype=float)
    bottom = N - p - 1
    return AIC(rss, p, N) + top / bottom


def likelihood(aiccA, aiccB):
    return -1.0 * (aiccA - aiccB) / 2.0


def expand_mask(mask, expand):
    selem = np.ones(expand, dtype=bool)
    selem.shape += (1, 1,)
    return nd.binary_dilation(mask, selem)


def get_rms(residual):
    diff = residual - np.roll(residual, 2, axis=0)
    return 1.4826 * np.nanmedian(np.abs(diff), axis=0) / 2 ** 0.5


def get_rss(cube, model, expand=20, usemask=True, mask=None, include_nosamp=True):
    """UltraCube.py:1384-1475 (planemask=None branch).  Mutates ``mask`` and ``model`` like the reference.
    Returns (rss, nsamp, final_mask_float)."""
    if usemask:
        if mask is None:
            mask = model > 0
    else:
        mask = ~np.isnan(model)
    if include_nosamp:
        nsamp_map = np.nansum(mask, axis=0)
        mm = nsamp_map <= 0
        try:
            max_y, max_x = np.where(nsamp_map == np.nanmax(nsamp_map))
            spec_mask_fill = copy(mask[:, max_y[0], max_x[0]])
        except Exception:
            spec_mask_fill = np.any(mask, axis=(1, 2))
        mask[:, mm] = spec_mask_fill[:, np.newaxis]
    model[np.isnan(model)] = 0
    if expand > 0:
        mask = expand_mask(mask, expand)
    mask = mask.astype(float)
    residual = cube - model
    rss = np.nansum((residual * mask) ** 2, axis=0)
    rss[rss == 0] = np.nan
    nsamp = np.nansum(mask, axis=0)
    nsamp[np.isnan(rss)] = np.nan
    return rss, nsamp, mask


def get_chisq(cube, model, expand=20, reduced=True, usemask=True, mask=None):
    if usemask:
        if mask is None:
            mask = model > 0
    else:
        mask = ~np.isnan(model)
    residual = cube - model
    if expand > 0:
        mask = expand_mask(mask, expand)
    mask = mask.astype(float)
    chisq = np.nansum((residual * mask) ** 2, axis=0)
    if reduced:
        reduction = np.nansum(mask, axis=0)
        reduction[reduction == 0] = np.nan
        chisq /= reduction
    rms = get_rms(residual)
    chisq /= rms ** 2
    if reduced:
        return chisq
    return chisq, np.nansum(mask, axis=0)


def rss_maps(cube, model, mask, expand=20):
    """UltraCube.get_rss method (:432-451): get_rss then nsamp<1 -> NaN, rss<=0 -> NaN."""
    rss, nsamp, _ = get_rss(cube, model, expand=expand, usemask=True, mask=mask)
    bad = nsamp < 1
    nsamp[bad] = np.nan
    bad = np.logical_or(bad, rss <= 0)
    rss[bad] = np.nan
    return rss, nsamp


def has_fit(parcube):
    mask = np.any(parcube != 0, axis=0)
    return mask & np.any(np.isfinite(parcube), axis=0)


def all_lnk_maps(cube, model1, model2, expand=20):
    """get_all_lnk_maps(ncomp_max=2) (:1218-1280) with master_model_mask rebuilt from both model cubes
    (reset_model_mask([2,1]) :359-368).  model1/model2 are the ``_modelcube`` arrays (NaN where no fit).
    Returns dict(lnk10, lnk20, lnk21, rss0, rss1, rss2, nsamp)."""
    with np.errstate(all="ignore"):
        mask = np.logical_or(model2 > 0, model1 > 0)          # NaN > 0 is False
        m1 = model1.copy()
        m2 = model2.copy()
        # evaluation order of the reference: lnk21 -> get_AICc(2), get_AICc(1); lnk20 -> get_AICc(0)
        # (cached 2); lnk10 (cached).  The shared mask is mutated in place by the first call
        # (include_nosamp fill), so later calls see the filled mask -- same result as filling once.
        rss2, n2 = rss_maps(cube, m2, mask, expand)
        rss1, n1 = rss_maps(cube, m1, mask, expand)
        rss0, n0 = rss_maps(cube, np.zeros(cube.shape), mask, expand)
        a2 = AICc(rss2, 8, n2)
        a1 = AICc(rss1, 4, n1)
        a0 = AICc(rss0, 0, n0)
        return dict(lnk10=likelihood(a1, a0), lnk20=likelihood(a2, a0), lnk21=likelihood(a2, a1),
                    rss0=rss0, rss1=rss1, rss2=rss2, nsamp=n2, nsamp1=n1, nsamp0=n0)


def best_2c_parcube(parcube1, errcube1, parcube2, errcube2, lnk10, lnk20, lnk21,
                    lnk21_thres=5, lnk20_thres=5, lnk10_thres=5):
    """get_best_2c_parcube (:1283-1379, include_1c=True, return_lnks=True) + the integer ncomp map."""
    with np.errstate(invalid="ignore"):
        parcube = parcube2.copy()
        errcube = errcube2.copy()
        mask = np.logical_and(lnk21 > lnk21_thres, lnk20 > lnk20_thres)
        parcube[:4, ~mask] = parcube1[:4, ~mask]
        errcube[:4, ~mask] = errcube1[:4, ~mask]
        parcube[4:8, ~mask] = np.nan
        errcube[4:8, ~mask] = np.nan
        ncomp = np.where(mask, 2, 1).astype(np.int8)
        mask1 = lnk10 > lnk10_thres
        parcube[:, ~mask1] = np.nan
        errcube[:, ~mask1] = np.nan
        ncomp[~mask1] = 0
        lnk10 = lnk10.copy()
        lnk20 = lnk20.copy()
        lnk21 = lnk21.copy()
        lnk10[~has_fit(parcube1)] = np.nan
        lnk20[~has_fit(parcube2)] = np.nan
        lnk21[~has_fit(parcube2)] = np.nan
    return parcube, errcube, lnk10, lnk20, lnk21, ncomp
