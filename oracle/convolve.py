"""Convolved-cube stage (oracle).  TEST INFRASTRUCTURE -- never imported by the product.

Restates, in float64 numpy, what ``convolve_tools.convolve_sky_byfactor`` (mufasa/convolve_tools.py:37-87) does to
the voxels of a cube through its third-party calls:
  * the smoothing kernel: ``Beam(factor*BMAJ, factor*BMIN, BPA)`` reached from the cube's own beam, i.e. the
    deconvolved beam ``sqrt(factor^2 - 1) * (BMAJ, BMIN)`` at the same angle, sampled as radio_beam's
    ``EllipticalGaussian2DKernel`` (Gaussian2D at integer pixel offsets, x_stddev along ``BPA + 90 deg`` from the
    x axis, support 8 sigma_major rounded up to an odd size)                                   :51-59, :121
  * ``SpectralCube.convolve_to`` = astropy ``convolve(plane, kernel, normalize_kernel=True)`` with its defaults
    boundary='fill' (zeros), nan_treatment='interpolate'; the result keeps the cube's mask          :121
  * ``reproject(header, order='bilinear')`` onto the header made by ``downsample_header``
    (mufasa/utils/fits_utils.py:11-60)                                                           :72-79
PARITY UNPINNED against the third-party packages themselves (radio_beam, spectral_cube, astropy.convolution and
reproject are absent from this image, the reference has no fixtures): their behaviour is [RECALLED].  What IS
pinned: ``downsample_header`` against the reference's own code (tests/golden/cnv_golden.npz), the convolution
against scipy.ndimage.correlate and the interpolation against scipy.ndimage.map_coordinates on NaN-free input
(tests/test_oracle_golden.py).
"""
import numpy as np

FWHM_TO_SIGMA = 1.0 / np.sqrt(8.0 * np.log(2.0))


def beam_kernel(bmaj, bmin, bpa_deg, pixscale, factor, support_scaling=8):
    """Kernel [K, K] (float64, unnormalised) that takes a (bmaj, bmin, bpa) beam to ``factor`` times itself.
    ``bmaj``, ``bmin``, ``pixscale`` in the same angular unit."""
    grow = np.sqrt(float(factor) ** 2 - 1.0)
    s_maj = grow * bmaj * FWHM_TO_SIGMA / pixscale
    s_min = grow * bmin * FWHM_TO_SIGMA / pixscale
    theta = np.radians(bpa_deg + 90.0)
    size = int(np.ceil(support_scaling * s_maj))
    if size % 2 == 0:
        size += 1
    r = np.arange(size) - size // 2
    x, y = np.meshgrid(r, r)
    ct, st = np.cos(theta), np.sin(theta)
    a = 0.5 * (ct ** 2 / s_maj ** 2 + st ** 2 / s_min ** 2)
    b = 0.5 * (np.sin(2 * theta) / s_maj ** 2 - np.sin(2 * theta) / s_min ** 2)
    c = 0.5 * (st ** 2 / s_maj ** 2 + ct ** 2 / s_min ** 2)
    return np.exp(-(a * x ** 2 + b * x * y + c * y ** 2)) / (2 * np.pi * s_maj * s_min)


def convolve_planes(cube, kernel, include=None):
    """astropy ``convolve`` on every plane of ``cube`` [nchan, ny, nx]; ``include`` [ny, nx] bool = the cube's 2-D
    mask (False -> NaN on input); voxels that were NaN / masked on input are NaN on output."""
    cube = np.asarray(cube, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    K = kernel.shape[0]
    h = K // 2
    nchan, ny, nx = cube.shape
    good = np.isfinite(cube)
    if include is not None:
        good &= np.asarray(include, bool)[None]
    val = np.pad(np.where(good, cube, 0.0), ((0, 0), (h, h), (h, h)))
    fin = np.pad(good.astype(np.float64), ((0, 0), (h, h), (h, h)), constant_values=1.0)   # fill zeros are finite
    num = np.zeros_like(cube)
    den = np.zeros_like(cube)
    for dy in range(K):
        for dx in range(K):
            w = kernel[dy, dx]
            num += w * val[:, dy:dy + ny, dx:dx + nx]
            den += w * fin[:, dy:dy + ny, dx:dx + nx]
    with np.errstate(all="ignore"):
        out = np.where(den > 0, num / den, np.nan)
    out[~good] = np.nan          # the convolved cube keeps the input cube's mask
    return out


def resample_bilinear(cube, out_ny, out_nx, ay, by, ax, bx):
    """Every plane sampled at (y, x) = (ay oy + by, ax ox + bx): bilinear, the edge replicated within half a pixel of
    the image, NaN beyond; a NaN neighbour gives NaN."""
    cube = np.asarray(cube, dtype=np.float64)
    nchan, ny, nx = cube.shape
    y = ay * np.arange(out_ny, dtype=np.float64) + by
    x = ax * np.arange(out_nx, dtype=np.float64) + bx
    oky = (y >= -0.5) & (y <= ny - 0.5)
    okx = (x >= -0.5) & (x <= nx - 0.5)
    y = np.clip(y, 0, ny - 1)
    x = np.clip(x, 0, nx - 1)
    yi = np.minimum(np.floor(y).astype(int), ny - 1)
    xi = np.minimum(np.floor(x).astype(int), nx - 1)
    yj = np.minimum(yi + 1, ny - 1)
    xj = np.minimum(xi + 1, nx - 1)
    fy = (y - yi)[None, :, None]
    fx = (x - xi)[None, None, :]
    v00 = cube[:, yi][:, :, xi]
    v01 = cube[:, yi][:, :, xj]
    v10 = cube[:, yj][:, :, xi]
    v11 = cube[:, yj][:, :, xj]
    out = (1 - fy) * ((1 - fx) * v00 + fx * v01) + fy * ((1 - fx) * v10 + fx * v11)
    out[:, ~oky, :] = np.nan
    out[:, :, ~okx] = np.nan
    return out


def downsample_header(header, factor, axis):
    """mufasa/utils/fits_utils.py:11-60 on a dict."""
    header = dict(header)
    cd, cp = 'CDELT%d' % axis, 'CRPIX%d' % axis
    scale = 1.0 / factor
    header[cp] = (header[cp] - 1) * scale + scale / 2.0 + 0.5
    header[cd] = header[cd] * factor
    return header


def downsample_map(hdr, factor):
    """Header of the regridded cube (convolve_tools.py:72-77) and the 0-based pixel map onto the input grid:
    in = a * out + b per axis, exact when both headers share the projection (same CRVAL / CTYPE)."""
    nh = downsample_header(downsample_header(hdr, factor, 1), factor, 2)
    nh['NAXIS1'] = int(np.rint(hdr['NAXIS1'] / factor))
    nh['NAXIS2'] = int(np.rint(hdr['NAXIS2'] / factor))
    coef = {}
    for ax in (1, 2):
        a = nh['CDELT%d' % ax] / hdr['CDELT%d' % ax]
        b = (hdr['CRPIX%d' % ax] - 1) - a * (nh['CRPIX%d' % ax] - 1)
        coef[ax] = (a, b)
    return nh, coef
