"""Fill NaN holes of [ny, nx] / [k, ny, nx] maps by growing Gaussian smoothing (mirror of
mufasa/utils/interpolate.py:5-47; called from cubefit_gen, multi_v_fit.py:290-291)."""
import numpy as np

from .convolve import convolve_gauss2d


def expand_interpolate(data, ker_sig=1):
    if data.ndim == 2:
        sm = convolve_gauss2d(data, ker_sig, boundary='extend')
        q = np.isnan(data)
        data[q] = sm[q]
    elif data.ndim == 3:
        for i in range(data.shape[0]):
            sm = convolve_gauss2d(data[i], ker_sig, boundary='extend')
            q = np.isnan(data[i])
            data[i][q] = sm[q]
    return data


def iter_expand(data, mask):
    """Smooth with kernels of sigma 1, 2, 3, ... until every pixel of ``mask`` is finite; NaN outside the mask."""
    def need(d):
        holes = np.any(np.isnan(d), axis=0) if d.ndim == 3 else np.isnan(d)
        return np.sum(holes[mask]) > 0

    i = 0
    while need(data):
        data = expand_interpolate(data, ker_sig=1 + i)
        i += 1
        if i > max(data.shape[-2:]):          # nothing finite anywhere: the reference would loop forever
            break
    if data.ndim == 3:
        data[:, ~mask] = np.nan
    else:
        data[~mask] = np.nan
    return data
