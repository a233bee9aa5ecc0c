"""Host-side helpers on 2-D MAPS (masks, parameter maps) that the drivers use around the fit: skimage-style binary
morphology (erosion with border_value=True, dilation, opening, remove_small_objects / remove_small_holes), mad_std,
edge trimming of a footprint (signals.trim_cube_edge :199-226 reduces to this on the 2-D plane), the histogram
mode (estimate_mode :399-420) and a Gaussian smoother.  [RECALLED semantics of skimage / astropy.stats]

Everything that touches the CUBE -- robust rms, S/N, signal masks, velocity estimates, moments (signals.get_moments
:329-396 and its callees) -- runs on the GPU (prepass.py, csrc/b200fit_prepass.cu); the numpy restatement those
kernels are tested against lives in oracle/signals.py, not here.
"""
import numpy as np
from scipy import ndimage as nd

from .utils.convolve import convolve_gauss2d  # noqa: F401  (re-exported for the drivers)

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


def remove_small_objects(mask, min_size=64):
    """skimage.morphology.remove_small_objects on a boolean map: drop 4-connected components smaller than min_size."""
    mask = np.asarray(mask, bool)
    lab, n = nd.label(mask)
    if n == 0:
        return mask.copy()
    sizes = np.bincount(lab.ravel())
    small = sizes < min_size
    small[0] = False
    return mask & ~small[lab]


def remove_small_holes(mask, area_threshold=64):
    """skimage.morphology.remove_small_holes: fill 4-connected background components smaller than the threshold."""
    mask = np.asarray(mask, bool)
    return ~remove_small_objects(~mask, area_threshold)


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


def estimate_mode(data, bins=25):
    data = data[np.isfinite(data)]
    counts, edges = np.histogram(data, bins=bins)
    k = np.argmax(counts)
    return (edges[k] + edges[k + 1]) / 2


