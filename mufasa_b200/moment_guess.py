"""Moment-based initial guesses: host mirror of mufasa/moment_guess.py (moment_guesses :388-518, LineSetup :27-58,
get_tex / get_tau / peakT :774-800).  Produces the per-pixel guesses the CUDA fit kernel stages."""
import numpy as np

# moment_guess.py:15 imports h, kb from pyspeckit's ammonia_constants; CODATA-2018 cgs values
h = 6.62607015e-27
kb = 1.380649e-16
TBG = 2.7315


class LineSetup(object):
    """Guessing ranges of Tex and tau per species (moment_guess.py:27-58)."""

    def __init__(self, linetype='nh3', new_recipe=True):
        if linetype not in ('nh3', 'n2hp'):
            raise Exception("{} is an invalid linetype".format(linetype))
        if new_recipe:
            self.tex_max, self.tau_max, self.tex_min, self.tau_min = 8.0, 5.0, 3.0, 0.1
        else:
            self.tex_max, self.tau_max, self.tex_min, self.tau_min = 8.0, 1.0, 3.1, 0.3


def moment_guesses(moment1, moment2, ncomp, sigmin=0.07, tex_guess=3.2, tau_guess=0.5, moment0=None, linetype='nh3',
                   mom0_floor=None):
    """[v, sigma, Tex, tau] x ncomp from moment maps.  Same recipe and the same quirks as the reference: ``moment2``
    is used as a width whatever it is, gs_sig clamps IN PLACE into moment2/ncomp, and the Tex/tau range clamps happen
    after the values were copied into the output (they do not affect it)."""
    ls = LineSetup(linetype)
    if moment0 is not None:
        norm_ref = 99.73
        mom0high = np.percentile(moment0[np.isfinite(moment0)], norm_ref)
        if mom0_floor is None:
            m0n = moment0.copy() * norm_ref / 100.0 / mom0high
            tex_guess = m0n * ls.tex_max
            tau_guess = m0n * ls.tau_max
        else:
            m0n = moment0.copy() * norm_ref / 100.0 / (mom0high - mom0_floor)
            tex_guess = m0n * (ls.tex_max - ls.tex_min) + ls.tex_min
            tau_guess = m0n * (ls.tau_max - ls.tau_min) + ls.tau_min
    m1, m2 = moment1, moment2
    gs_sig = m2 / ncomp
    with np.errstate(invalid="ignore"):
        gs_sig[gs_sig < sigmin] = sigmin
    gg = np.zeros((ncomp * 4,) + m1.shape)
    if ncomp == 1:
        gg[0], gg[1], gg[2], gg[3] = m1, gs_sig, tex_guess, tau_guess
    elif ncomp == 2:
        sigmaoff, tau2_frac = 0.4, 0.25
        gg[0] = m1 - sigmaoff * m2
        gg[1] = gs_sig
        gg[2] = tex_guess
        gg[3] = tau_guess * (1 - tau2_frac)
        gg[4] = m1 + sigmaoff * m2
        gg[5] = gs_sig
        gg[6] = tex_guess * 0.8
        gg[7] = tau_guess * tau2_frac
    else:
        raise NotImplementedError("the CUDA fitter covers 1- and 2-component models")
    return gg


def get_tex(Ta, tau=0.5, nu=23.722634):
    T0 = (h * nu * 1e9 / kb)
    term1 = Ta / T0 / (1 - np.exp(-tau))
    term2 = 1.0 / (np.exp(T0 / TBG) - 1)
    return T0 / np.log(1.0 / (term1 + term2) + 1)


def get_tau(Ta, tex=6.0, nu=23.722634):
    T0 = (h * nu * 1e9 / kb)
    term1 = 1 / (np.exp(T0 / tex) - 1) - 1 / (np.exp(T0 / TBG) - 1)
    return -np.log(-Ta / T0 / term1 + 1)


def peakT(tex, tau, nu=23.722634):
    T0 = (h * nu * 1e9 / kb)
    return (T0 / (np.exp(T0 / tex) - 1) - T0 / (np.exp(T0 / TBG) - 1)) * (1 - np.exp(-tau))


# ---- windowed moments and the guesses for widely separated components ------------------------------------------
# moment_guess.py:119-367, 523-701.  The reference takes these moments with spectral_cube (``vmask_moments``) or
# with pyspeckit's ``Cube.momenteach`` (``window_moments`` on a pyspeckit.Cube); both are absent from the image
# [RECALLED, parity unpinned]:
#   spectral_cube   moment0 = sum d dv,  moment1 = sum v d / sum d,  moment2 = sum (v - m1)^2 d / sum d
#   pyspeckit       spectrum.moments.moments(xarr, data, vheight=False) -> [amplitude, x, width]:  amplitude = the
#                   maximum of the data (the minimum when |min| > |max|), x = sum v d / sum d,
#                   width = sqrt(|sum (v - x)^2 d / sum d|); NaN channels are masked out (Spectrum.data is a masked
#                   array); pixels outside the cube's maskmap keep the zeros momentcube starts from.
# They are row reductions over the whole cube: torch tensor ops on the GPU when there is one (plumbing around the
# kernels, off the fit path -- as the product maps of master_fitter._best_model_products).
def _torch_device(device=None):
    import torch
    if device is not None and torch.cuda.is_available():
        return torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _nanmedian0(d):
    """numpy's nanmedian along axis 0 of a [n, ...] torch tensor (mean of the two middle values of an even count)."""
    import torch
    fin = torch.isfinite(d)
    n = fin.sum(dim=0, keepdim=True)
    srt = torch.sort(torch.where(fin, d, torch.full_like(d, float("inf"))), dim=0).values
    lo, hi = torch.clamp((n - 1) // 2, min=0), torch.clamp(n // 2, min=0)
    med = 0.5 * (torch.gather(srt, 0, lo) + torch.gather(srt, 0, hi))
    return torch.where(n > 0, med, torch.full_like(med, float("nan")))[0]


def mad_std_cube(data, device=None):
    """astropy ``mad_std(cube, axis=0, ignore_nan=True)``: 1.4826 median|x - median x| per pixel, on the GPU when there
    is one (``signals.mad_std`` is its numpy twin; at 512 x 512 x 1000 numpy's two nanmedians take ~15 s)."""
    import torch
    d = torch.as_tensor(np.ascontiguousarray(data), device=_torch_device(device)).to(torch.float64)
    med = _nanmedian0(d)
    return (1.482602218505602 * _nanmedian0((d - med[None]).abs())).cpu().numpy()


def _window_include(data, v, vmid, hwidth):
    """[nchan, ny, nx] bool: finite voxels with vmid - hwidth < v < vmid + hwidth (window_mask_pcube :756-768)."""
    import torch
    inc = torch.isfinite(data)
    if vmid is not None:
        vm = torch.as_tensor(np.asarray(vmid, dtype=np.float64), device=data.device)
        vv = v[:, None, None]
        inc &= (vv > vm - hwidth) & (vv < vm + hwidth)     # a NaN centre excludes the pixel
    return inc


def pys_moments(cube_data, v_kms, include=None, vmid=None, hwidth=None, maskmap=None, device=None):
    """pyspeckit ``momenteach(vheight=False)`` maps (amplitude, x, width) [3, ny, nx] of the channels in ``include``
    (and/or inside the window ``vmid`` +- ``hwidth``); 0 where a pixel has no such channel or is outside ``maskmap``."""
    import torch
    dev = _torch_device(device)
    d = torch.as_tensor(np.ascontiguousarray(cube_data), device=dev).to(torch.float64)
    v = torch.as_tensor(np.asarray(v_kms, dtype=np.float64), device=dev)
    inc = _window_include(d, v, vmid, hwidth)
    if include is not None:
        inc &= torch.as_tensor(np.ascontiguousarray(include), device=dev)
    zero = torch.zeros((), dtype=torch.float64, device=dev)
    dz = torch.where(inc, d, zero)
    tot = dz.sum(0)
    x = (dz * v[:, None, None]).sum(0) / tot
    width = torch.sqrt(torch.abs((dz * (v[:, None, None] - x) ** 2).sum(0) / tot))
    dmax = torch.where(inc, d, torch.full_like(d, -float("inf"))).amax(0)
    dmin = torch.where(inc, d, torch.full_like(d, float("inf"))).amin(0)
    amp = torch.where(dmin.abs() > dmax.abs(), dmin, dmax)
    valid = inc.any(0)
    if maskmap is not None:
        valid &= torch.as_tensor(np.asarray(maskmap, bool), device=dev)
    out = torch.stack([amp, x, width])
    out = torch.where(valid[None], out, torch.zeros_like(out))
    return out.cpu().numpy()


def window_mask_pcube(cube, vmid, win_hwidth=4.0):
    """moment_guess.py:756-768 on a SpecCube: (data with the channels outside the window set to NaN, maskmap)."""
    v = np.asarray(cube.spectral_axis, dtype=np.float64)[:, None, None]
    with np.errstate(invalid="ignore"):
        smask = np.logical_and(v > vmid - win_hwidth, v < vmid + win_hwidth)
    data = np.where(smask, cube._data, np.nan).astype(cube._data.dtype)
    return data, np.any(np.isfinite(data), axis=0)


def window_moments(cube, window_hwidth=4.0, v_atpeak=None, maskmap=None, device=None):
    """moment_guess.py:140-367 for a cube handed over as pyspeckit.Cube (``moments_pys_cube``): pyspeckit moments
    inside a per-pixel window about ``v_atpeak``; without ``v_atpeak`` a first pass over the whole spectrum supplies
    it.  Returns (m0 = amplitude, m1, m2 = width) maps."""
    v = cube.spectral_axis
    if v_atpeak is None:
        v_atpeak = pys_moments(cube._data, v, maskmap=maskmap, device=device)[1]
    m = pys_moments(cube._data, v, vmid=v_atpeak, hwidth=window_hwidth, maskmap=maskmap, device=device)
    return m[0], m[1], m[2]


def vmask_cube(cube, vmap, window_hwidth=3.0):
    """moment_guess.py:130-137: [nchan, ny, nx] bool of the channels within ``window_hwidth`` of ``vmap``."""
    v = np.asarray(cube.spectral_axis, dtype=np.float64)[:, None, None]
    with np.errstate(invalid="ignore"):
        return np.logical_and(v > vmap - window_hwidth, v < vmap + window_hwidth)


def vmask_moments(cube, vmap, window_hwidth=3.0, device=None):
    """moment_guess.py:119-127: spectral_cube moments 0 / 1 / sqrt|2| inside the per-pixel window about ``vmap``
    (NaN where a pixel has no finite channel in its window)."""
    import torch
    dev = _torch_device(device)
    d = torch.as_tensor(np.ascontiguousarray(cube._data), device=dev).to(torch.float64)
    v = torch.as_tensor(np.asarray(cube.spectral_axis, dtype=np.float64), device=dev)
    inc = _window_include(d, v, vmap, window_hwidth)
    if cube.plane_include is not None:
        inc &= torch.as_tensor(np.asarray(cube.plane_include, bool), device=dev)[None]
    dz = torch.where(inc, d, torch.zeros((), dtype=torch.float64, device=dev))
    dv = abs(float(cube.spectral_axis[1] - cube.spectral_axis[0]))
    tot = dz.sum(0)
    m1 = (dz * v[:, None, None]).sum(0) / tot
    m2 = torch.sqrt(torch.abs((dz * (v[:, None, None] - m1) ** 2).sum(0) / tot))
    m0 = torch.where(inc.any(0), tot * dv, torch.full_like(tot, float("nan")))
    return m0.cpu().numpy(), m1.cpu().numpy(), m2.cpu().numpy()


def moment_guesses_1c(m0, m1, m2):
    """moment_guess.py:523-572: [v, sigma, Tex, tau] from pyspeckit-style moments -- ``m0`` is an amplitude, turned
    into an integrated flux with the width (the reference takes the square root of the width once more: kept); above
    3 K km/s tau is fixed at 2.5 and Tex follows, below it Tex is fixed at 6 K and tau follows."""
    with np.errstate(invalid="ignore", divide="ignore"):
        gs_sig = np.asarray(m2, dtype=np.float64) ** 0.5
        mom0 = np.asarray(m0, dtype=np.float64) * gs_sig * np.sqrt(2 * np.pi)
        mom0_thres, mom0_min, tau_fx, tex_fx = 3.0, 0.03, 2.5, 6.0
        if mom0.ndim == 0:
            mom0 = max(float(mom0), mom0_min) if mom0 == mom0 else float(mom0)
            if mom0 > mom0_thres:
                tau_guess, tex_guess = tau_fx, get_tex(mom0, tau_fx)
            else:
                tex_guess, tau_guess = tex_fx, get_tau(mom0, tex_fx)
        elif mom0.ndim == 2:
            mom0[mom0 < mom0_min] = mom0_min
            tau_guess, tex_guess = np.zeros(mom0.shape), np.zeros(mom0.shape)
            mask = mom0 > mom0_thres
            tau_guess[mask] = tau_fx
            tex_guess[mask] = get_tex(mom0[mask], tau_fx)
            tau_guess[~mask] = get_tau(mom0[~mask], tex_fx)
            tex_guess[~mask] = tex_fx
        else:
            raise Exception("the moment 0 input has the wrong dimension ({})".format(mom0.ndim))
    return np.asarray([m1, gs_sig, tex_guess, tau_guess])


def mom_guess_wide_sep(cube, vpeak=None, rms=None, planemask=None, multicore=None, device=None):
    """moment_guess.py:575-701 (the SpectralCube branch): two-component guesses for widely separated components.
    Component 1 from the windowed moments about the spectrum's centroid; the emission within 2 sigma of it and
    everything further than 4 km/s is blanked, and the moments of what is left give component 2.  Where that
    remainder is fainter than 2 rms both components start from component 1, one width apart, with half its tau."""
    win_hwidth, f_sig_mask, rms_thres, f_tau, f_sig = 4.0, 2.0, 2.0, 0.5, 0.5
    if rms is None:
        rms = mad_std_cube(cube._data, device=device)
    v = np.asarray(cube.spectral_axis, dtype=np.float64)
    m0, m1, m2 = window_moments(cube, window_hwidth=win_hwidth, v_atpeak=vpeak, maskmap=planemask, device=device)
    gg1 = moment_guesses_1c(m0, m1, m2)
    with np.errstate(invalid="ignore"):
        sig = m2 ** 0.5
        vv = v[:, None, None]
        smask = np.logical_and(vv > m1 - sig * f_sig_mask, vv < m1 + sig * f_sig_mask)
        smask2 = np.logical_and(vv > m1 - win_hwidth, vv < m1 + win_hwidth)
        smask = np.logical_or(smask, ~smask2)
    # with_mask(~smask).with_fill_value(0): masked voxels read as zeros, and so do the cube's own NaNs
    data2 = np.where(smask | ~np.isfinite(cube._data), 0.0, cube._data).astype(np.float32)
    m0n, m1n, m2n = pys_moments(data2, v, maskmap=planemask, device=device)
    gg2 = moment_guesses_1c(m0n, m1n, m2n)
    with np.errstate(invalid="ignore"):
        mask = m0n < rms * rms_thres
    gg2[:, mask] = gg1[:, mask].copy()
    gg = np.concatenate((gg1, gg2), axis=0)
    gg[0, mask] += sig[mask]
    gg[4, mask] -= sig[mask]
    gg[3, mask] *= f_tau
    gg[7, mask] *= f_tau
    gg[1, :] *= f_sig
    gg[5, :] *= f_sig
    return gg
