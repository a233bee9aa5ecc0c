"""Guess refinement on the host: mirror of mufasa/guess_refine.py -- tautex_renorm :183-213 (from cubefit_gen,
multi_v_fit.py:330-331) and the functions that turn the fit of the CONVOLVED cube into guesses for the full-resolution
fit (guess_from_cnvpara :113-180 with quick_2comp_sort :36-102, mask_swap_2comp :105-110, simple_para_clean :254-275,
refine_each_comp :216-251, refine_guess :317-393, master_mask / mask_cleaning :286-298, refine_2c_guess :395-416).
These work on [k, ny, nx] maps of the DOWNSAMPLED grid (cfg3: 256 x 256): small host arrays, numpy / scipy.
skimage / astropy calls are replaced by the stand-ins of signals.py and utils/convolve.py.
"""
import numpy as np
from scipy.ndimage import median_filter

from . import moment_guess as mmg
from . import signals
from .utils import interpolate
from .utils.convolve import convolve_gauss2d

nu0_nh3 = 23.6944955


def tautex_renorm(taumap, texmap, tau_thresh=0.21, tex_thresh=15.0, nu=nu0_nh3):
    """Re-normalise the degenerate (tau, Tex) pair of optically thin guesses: tau below ``tau_thresh`` is raised to
    the threshold and Tex lowered so the peak brightness is kept.  In place; returns (taumap, texmap).

    Reference quirks kept: get_tex is called WITHOUT nu (its default 23.722634 GHz applies, guess_refine.py:199), and
    the reference's second, high-Tex branch assigns through chained fancy indexing (:209-210) and so changes
    nothing -- it is omitted here."""
    f_tau = 0.5
    isthin = np.logical_and(taumap < tau_thresh, np.isfinite(taumap))
    TA_lowtau = mmg.peakT(texmap[isthin], taumap[isthin] * f_tau, nu=nu)
    texmap[isthin] = mmg.get_tex(TA_lowtau, tau=tau_thresh * f_tau)
    taumap[isthin] = tau_thresh
    return taumap, texmap


def mask_swap_2comp(data_cnv, swapmask):
    """Exchange component 1 and 2 (parameters 0-3 <-> 4-7, errors 8-11 <-> 12-15) where ``swapmask``."""
    out = data_cnv.copy()
    for a, b in ((0, 4), (8, 12)):
        out[a:a + 4, swapmask] = data_cnv[b:b + 4, swapmask]
        out[b:b + 4, swapmask] = data_cnv[a:a + 4, swapmask]
    return out


def quick_2comp_sort(data_cnv, filtsize=2, method="error_v", nu=nu0_nh3, f_tau=0.5):
    """Order the two components of a [16, ny, nx] parameter + error stack."""
    with np.errstate(invalid="ignore"):
        if method == "error_v":          # the component with the smaller vlsr error first
            data_cnv = mask_swap_2comp(data_cnv, data_cnv[8] > data_cnv[12])
        elif method == "Tpeak":          # reference indices kept (3, 4 / 6, 7: guess_refine.py:49-50)
            Ta = mmg.peakT(data_cnv[3], data_cnv[4] * f_tau, nu=nu)
            Tb = mmg.peakT(data_cnv[6], data_cnv[7] * f_tau, nu=nu)
            data_cnv = mask_swap_2comp(data_cnv, Tb > Ta)
        elif method == "tautex":
            data_cnv = mask_swap_2comp(data_cnv, data_cnv[6] * data_cnv[7] > data_cnv[3] * data_cnv[4])
        elif method == "tau":
            data_cnv = mask_swap_2comp(data_cnv, data_cnv[7] > data_cnv[4])
        elif method == "chen2020":
            data_cnv = mask_swap_2comp(data_cnv, data_cnv[8] > data_cnv[12])
            ref = median_filter(data_cnv[8], size=(filtsize, filtsize))
            data_cnv = mask_swap_2comp(data_cnv, np.abs(data_cnv[8] - ref) > np.abs(data_cnv[12] - ref))

            def dist(p1, p2):
                pref = median_filter(p1, size=(filtsize, filtsize))
                return np.abs(p1 - pref), np.abs(p2 - pref)

            va, vb = dist(data_cnv[0], data_cnv[4])
            sa, sb = dist(data_cnv[1], data_cnv[5])
            data_cnv = mask_swap_2comp(data_cnv, np.hypot(va, sa) > np.hypot(vb, sb))
    return data_cnv


def simple_para_clean(pmaps, ncomp, npara=4, std_thres=2):
    """Blank the components whose vlsr error lies more than ``std_thres`` robust sigmas above the median."""
    pmaps = pmaps.copy()
    pmaps[pmaps == 0] = np.nan
    for i in range(ncomp):
        verr = pmaps[(i + ncomp) * npara]
        good = verr[np.isfinite(verr)]
        std_vErr = signals.mad_std(good, axis=None)
        median_vErr = np.median(good)
        with np.errstate(invalid="ignore"):
            mask = verr > median_vErr + std_vErr * std_thres
        pmaps[i * npara:(i + 1) * npara, mask] = np.nan
        pmaps[(i + ncomp) * npara:(i + ncomp + 1) * npara, mask] = np.nan
    return pmaps


def mask_cleaning(mask):
    mask = signals.binary_dilation(mask, signals.disk(1))
    return signals.remove_small_holes(mask, 9)


def master_mask(pcube):
    return mask_cleaning(np.any(np.isfinite(pcube), axis=0))


def refine_guess(map, min=None, max=None, mask=None, disksize=1, scipy_interpolate=False):
    """Clip a parameter map to [min, max], then fill it over ``mask`` by NaN-interpolating smoothing
    (FWHM 2.5 pixel Gaussian, then iter_expand); the original finite values are kept."""
    map = map.copy()
    with np.errstate(invalid="ignore"):
        if min is not None:
            map[map < min] = np.nan
        if max is not None:
            map[map > max] = np.nan
    finite = np.isfinite(map)
    n_valid = np.sum(finite)
    if n_valid < 2:
        if n_valid == 1:
            map[:] = map[finite][0]
        elif min is not None:
            map[:] = min if max is None else (max + min) / 2
        elif max is not None:
            map[:] = max
        else:
            map[:] = 0.0
        return map
    if mask is None:
        mask = mask_cleaning(finite)
    zi = convolve_gauss2d(map, 2.5 / 2.355, boundary='extend')
    zi[finite] = map[finite]
    return interpolate.iter_expand(zi, mask=mask)


def refine_each_comp(guess_comp, mask=None, v_range=None, sig_range=None, tau_thresh=0.1):
    """Refine the [4, ny, nx] guesses of one component; Tex and tau get tighter limits than the fit itself."""
    Tex_min, Tex_max, Tau_min, Tau_max, disksize = 3.0, 8.0, 0.2, 8.0, 1.0
    if mask is None:
        mask = master_mask(guess_comp)
    vmin, vmax = (None, None) if v_range is None else v_range
    sigmin, sigmax = (None, None) if sig_range is None else sig_range
    guess_comp[0] = refine_guess(guess_comp[0], min=vmin, max=vmax, mask=mask, disksize=disksize)
    guess_comp[1] = refine_guess(guess_comp[1], min=sigmin, max=sigmax, mask=mask, disksize=disksize)
    guess_comp[3], guess_comp[2] = tautex_renorm(guess_comp[3], guess_comp[2], tau_thresh=tau_thresh)
    guess_comp[2] = refine_guess(guess_comp[2], min=Tex_min, max=Tex_max, mask=mask, disksize=disksize)
    guess_comp[3] = refine_guess(guess_comp[3], min=Tau_min, max=Tau_max, mask=mask, disksize=disksize)
    return guess_comp


def guess_from_cnvpara(data_cnv, header_cnv, header_target, mask=None, tau_thresh=1, clean_map=True):
    """Guesses [4 ncomp, ny, nx] on the grid of ``header_target`` from the parameter + error maps
    [8 ncomp, ny_c, nx_c] fitted to the convolved cube."""
    from .convolve_tools import get_celestial_hdr, regrid_many
    npara = 4
    ncomp = int(data_cnv.shape[0] / npara / 2)
    data_cnv = data_cnv.copy()
    data_cnv[data_cnv == 0] = np.nan
    std_thres = 3 if ncomp == 1 else 1
    if ncomp == 2:
        data_cnv = quick_2comp_sort(data_cnv, filtsize=2, method="error_v")
    if clean_map:
        data_cnv = simple_para_clean(data_cnv, ncomp, npara=npara, std_thres=std_thres)
    data_cnv = data_cnv[0:npara * ncomp]
    for i in range(ncomp):
        data_cnv[i * npara:(i + 1) * npara] = refine_each_comp(data_cnv[i * npara:(i + 1) * npara], mask,
                                                               tau_thresh=tau_thresh)
    hdr_conv, hdr_final = get_celestial_hdr(header_cnv), get_celestial_hdr(header_target)
    newmask = signals.remove_small_holes(np.any(np.isfinite(data_cnv), axis=0), 25)
    # per map: fill the footprint, regrid (cubic griddata), grow the footprint back by two pixels.  The regridding
    # of all maps shares one triangulation (convolve_tools.regrid_many; same values as map-by-map calls)
    filled = np.array([interpolate.iter_expand(gss, mask=newmask) for gss in data_cnv])
    guesses_final = []
    for g in regrid_many(filled, hdr_conv, hdr_final):
        mask_l = signals.binary_dilation(np.isfinite(g), signals.disk(2))
        guesses_final.append(interpolate.iter_expand(g, mask=mask_l))
    return np.array(guesses_final)


def refine_2c_guess(guesses, f_sigv=0.5):
    """Push the two velocity guesses apart by f_sigv x half their summed widths and halve the width guesses."""
    v1, v2, s1, s2 = guesses[0], guesses[4], guesses[1], guesses[5]
    vdiff = v1 - v2
    with np.errstate(invalid="ignore", divide="ignore"):
        sign = vdiff / np.abs(vdiff)
    shift = (s1 + s2) / 2 * f_sigv * sign
    guesses[0], guesses[4] = v1 + shift, v2 - shift
    guesses[1], guesses[5] = s1 / 2, s2 / 2
    return guesses
