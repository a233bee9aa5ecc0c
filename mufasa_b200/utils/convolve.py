"""NaN-interpolating Gaussian smoothing of [ny, nx] maps.

Stands in for ``astropy.convolution.convolve(map, Gaussian2DKernel(sigma), boundary=...)`` as the reference uses it
(multi_v_fit.py:300-306, :697-699; utils/interpolate.py:5-24; guess_refine.py): kernel support 8*sigma rounded up
to an odd size, kernel normalised to unit sum, NaN pixels ignored and the weights re-normalised
(nan_treatment='interpolate'), boundary 'fill' (zeros, astropy's default) or 'extend' (edge replication).
astropy is absent from this image; the semantics are restated from its documented behaviour.  [RECALLED]
"""
import numpy as np
from scipy import ndimage as nd


def gaussian2d_kernel(sigma):
    size = int(np.ceil(8 * sigma))
    if size % 2 == 0:
        size += 1
    r = np.arange(size) - size // 2
    g = np.exp(-0.5 * (r / float(sigma)) ** 2)
    k = np.outer(g, g) / (2 * np.pi * sigma ** 2)
    return k / k.sum()


def convolve_gauss2d(img, sigma, boundary='fill'):
    img = np.asarray(img, dtype=np.float64)
    k = gaussian2d_kernel(sigma)
    h = k.shape[0] // 2
    if boundary == 'extend':
        pad = np.pad(img, h, mode='edge')
    elif boundary in ('fill', None):
        pad = np.pad(img, h, mode='constant', constant_values=0.0)
    else:
        raise ValueError("boundary must be 'fill' or 'extend'")
    nan = np.isnan(pad)
    num = nd.correlate(np.where(nan, 0.0, pad), k, mode='constant', cval=0.0)
    den = nd.correlate((~nan).astype(np.float64), k, mode='constant', cval=0.0)
    with np.errstate(all="ignore"):
        out = np.where(den > 1e-12, num / den, np.nan)   # fully-NaN neighbourhood stays NaN
    return out[h:-h, h:-h]
