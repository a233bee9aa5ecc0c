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


def gaussian1d_kernel(sigma):
    """The 1-D factor of ``gaussian2d_kernel``: the normalised 2-D kernel is outer(g, g)."""
    size = int(np.ceil(8 * sigma))
    if size % 2 == 0:
        size += 1
    r = np.arange(size) - size // 2
    g = np.exp(-0.5 * (r / float(sigma)) ** 2)
    return g / g.sum()


def convolve_gauss2d(img, sigma, boundary='fill'):
    """Two 1-D passes per accumulator (the Gaussian kernel is separable; a K x K correlation of a 512 x 512 map costs
    60 ms, and the guess refinement of one pipeline stage calls this ~80 times with growing K)."""
    img = np.asarray(img, dtype=np.float64)
    g = gaussian1d_kernel(sigma)
    h = g.size // 2
    if boundary == 'extend':
        pad = np.pad(img, h, mode='edge')
    elif boundary in ('fill', None):
        pad = np.pad(img, h, mode='constant', constant_values=0.0)
    else:
        raise ValueError("boundary must be 'fill' or 'extend'")
    nan = np.isnan(pad)

    def sep(a):
        return nd.correlate1d(nd.correlate1d(a, g, axis=0, mode='constant', cval=0.0), g, axis=1, mode='constant',
                              cval=0.0)

    num = sep(np.where(nan, 0.0, pad))
    den = sep((~nan).astype(np.float64))
    with np.errstate(all="ignore"):
        out = np.where(den > 1e-12, num / den, np.nan)   # fully-NaN neighbourhood stays NaN
    return out[h:-h, h:-h]
