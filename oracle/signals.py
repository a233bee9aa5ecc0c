"""Pre-fit signal statistics (oracle): numpy/scipy restatement of mufasa/signals.py on whole cubes.  TEST INFRASTRUCTURE.

  get_moments :329-396, get_rms_robust :48-91, refine_rms :94-121, get_snr :24-45, trim_cube_edge :199-226,
  v_estimate :251-272, get_v_mask :296-326, estimate_mode :399-420.

The reference does this with spectral_cube masked cubes and skimage morphology; neither is installed here, so the
same operations are restated on plain arrays: a "masked cube" is (data[nchan, ny, nx], include[nchan, ny, nx] bool),
moments follow spectral_cube's definitions (m0 = sum(I)|dv|, m1 = sum(I v)/sum(I), m2 = sum(I (v - m1)^2)/sum(I),
all-masked -> NaN), and skimage's binary_erosion / binary_dilation / binary_opening map onto scipy.ndimage with
skimage's border conventions (erosion: border_value=True).  [RECALLED semantics of the two libraries -- parity
unpinned against spectral_cube / skimage themselves; the 2-D helpers and get_rms_robust's arithmetic are plain
numpy / scipy calls.]

The product computes all of this on the GPU (mufasa_b200/prepass.py + csrc/b200fit_prepass.cu); this file is what
tests/test_gpu_prepass.py and tests/test_gpu_pipeline.py compare those kernels with.  Nothing under mufasa_b200/
imports it.
"""
import numpy as np
from scipy import ndimage as nd

MAD_TO_STD = 1.482602218505602


def _data(cube):
    return cube._data if hasattr(cube, '_data') else np.asarray(cube)


def _vaxis(cube):
    return np.asarray(cube.spectral_axis, dtype=np.float64)


def disk(r):
    y, x = np.mgrid[-r:r + 1, -r:r + 1]
    return (x * x + y * y) <= r * r


def ball(r):
    z, y, x = np.mgrid[-r:r + 1, -r:r + 1, -r:r + 1]
    return (x * x + y * y + z * z) <= r * r


def binary_erosion(mask, footprint=None):
    if footprint is None:
        footprint = nd.generate_binary_structure(mask.ndim, 1)
    return nd.binary_erosion(mask, structure=footprint, border_value=True)


def binary_dilation(mask, footprint=None):
    if footprint is None:
        footprint = nd.generate_binary_structure(mask.ndim, 1)
    return nd.binary_dilation(mask, structure=footprint)


def binary_opening(mask, footprint=None):
    return binary_dilation(binary_erosion(mask, footprint), footprint)


def mad_std(data, axis=0, include=None):
    """1.4826 * median(|x - median(x)|) over finite (and included) samples."""
    d = np.asarray(data, dtype=np.float64)
    if include is not None:
        d = np.where(include, d, np.nan)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            med = np.nanmedian(d, axis=axis, keepdims=True)
            return MAD_TO_STD * np.nanmedian(np.abs(d - med), axis=axis)


def trim_edge(mask, trim=3):
    mask[[0, -1], :] = False
    mask[:, [0, -1]] = False
    return binary_erosion(mask, disk(trim))


def trim_cube_edge(cube, trim=3, min_size=1000):
    """Returns the include-mask [nchan, ny, nx] of the edge-trimmed cube (signals.py:199-226)."""
    mask = np.isfinite(_data(cube))
    planemask = np.any(mask, axis=0)
    if np.sum(planemask) > min_size:
        planemask = trim_edge(planemask, trim=trim)
        mask[:, ~planemask] = False
    return mask


def refine_rms(data, include, rms, sigma_cut=3, expand=20):
    with np.errstate(invalid="ignore"):
        sig_mask = data > sigma_cut * rms                     # LazyComparisonMask on the raw data
    sig_mask = binary_dilation(sig_mask, np.ones((expand, 1, 1), dtype=bool))
    rms = mad_std(data, axis=0, include=~sig_mask & include)
    return rms, sig_mask


def get_rms_robust(cube, sigma_cut=2.5, expand=20, include=None, return_sigmask=False):
    data = _data(cube)
    if include is None:
        include = np.isfinite(data)
    rms = mad_std(data, axis=0, include=include)
    rms, sig_mask = refine_rms(data, include, rms, sigma_cut=sigma_cut, expand=expand)
    return (rms, sig_mask) if return_sigmask else rms


def get_snr(cube, rms=None, include=None):
    data = _data(cube)
    if include is None:
        include = np.isfinite(data)
    if rms is None:
        rms = get_rms_robust(cube, include=include)
    import warnings
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmax(np.where(include, data, np.nan), axis=0) / rms


def get_v_at_peak(data, include, v):
    d = np.where(include, data, -np.inf)
    idx = np.argmax(d, axis=0)
    v_est = v[idx].astype(np.float64)
    v_est[~np.any(include, axis=0)] = np.nan
    return v_est


def v_estimate(data, include, v, rms, snr_cut=3):
    with np.errstate(invalid="ignore"):
        sig_mask = data > rms * snr_cut
    sig_mask = binary_erosion(sig_mask, ball(2))
    return get_v_at_peak(data, sig_mask & include, v)


def get_v_mask(data, include, v, v_center, rms=None, window_hwidth=5):
    snr_cut = 3
    vv = v[:, None, None]
    with np.errstate(invalid="ignore"):
        window_mask = (vv < (v_center + window_hwidth)) & (vv > (v_center - window_hwidth))
        if rms is not None:
            return binary_opening(data > rms * snr_cut) & window_mask
    return include & window_mask


def moments(data, include, v):
    """spectral_cube moment0 / moment1 / moment2 of the masked cube (all-masked pixel -> NaN)."""
    w = np.where(include & np.isfinite(data), data, 0.0).astype(np.float64)
    anyv = np.any(include & np.isfinite(data), axis=0)
    dv = abs(v[1] - v[0])
    s0 = w.sum(axis=0)
    vv = v[:, None, None]
    with np.errstate(all="ignore"):
        m0 = np.where(anyv, s0 * dv, np.nan)
        m1 = np.where(anyv, (w * vv).sum(axis=0) / s0, np.nan)
        m2 = np.where(anyv, (w * (vv - m1) ** 2).sum(axis=0) / s0, np.nan)
    return m0, m1, m2


def get_moments(cube, window_hwidth=5, linewidth_sigma=True, trim=3, return_rms=False, central_win_hwidth=None,
                include=None, rms=None):
    """signals.py:329-396 on (data, include).  ``include`` = mask of an already trimmed cube (cubefit_gen trims first
    and passes trim=None)."""
    snr_cut = 3
    noise_mask_cut = 2.5
    data = _data(cube)
    v = _vaxis(cube)
    if include is None:
        include = trim_cube_edge(cube, trim=trim) if trim is not None else np.isfinite(data)
    if rms is None:      # ``rms``: the same map computed per pixel on the GPU (PCube.rms_robust)
        rms = get_rms_robust(cube, sigma_cut=noise_mask_cut, include=include)
    v_peak = v_estimate(data, include, v, rms, snr_cut=snr_cut)
    if central_win_hwidth is not None:
        v_mode = estimate_mode(v_peak, bins=25)
        tmp = get_v_mask(data, include, v, v_center=v_mode, rms=rms, window_hwidth=central_win_hwidth)
        v_peak = v_estimate(data, tmp & include, v, rms, snr_cut=snr_cut)
    signal_mask = get_v_mask(data, include, v, v_peak, rms=rms, window_hwidth=window_hwidth)
    signal_mask = binary_dilation(signal_mask, disk(3)[np.newaxis, :])
    m0, m1, m2 = moments(data, signal_mask & include, v)
    if linewidth_sigma:
        with np.errstate(invalid="ignore"):
            m2 = np.sqrt(m2)
    out = (m0, m1, m2)
    if return_rms:
        out += (rms,)
    return out


def estimate_mode(data, bins=25):
    data = data[np.isfinite(data)]
    counts, edges = np.histogram(data, bins=bins)
    k = np.argmax(counts)
    return (edges[k] + edges[k + 1]) / 2
