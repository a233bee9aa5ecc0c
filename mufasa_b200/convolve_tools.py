"""Convolved-cube stage on the GPU: mirror of mufasa/convolve_tools.py (convolve_sky_byfactor :37-87, convolve_sky
:90-138, snr_mask :141-178, edge_trim :181-189, regrid :234-247, get_celestial_hdr :250-255).

The reference smooths the cube to ``factor`` x its beam with spectral_cube / radio_beam / astropy.convolution and
regrids it with reproject, all on the host (cfg3: 262 M voxels x a 23 x 23 kernel).  Here the cube is uploaded once
as channel-major planes; ``b200fit_convolve_planes`` and ``b200fit_resample_bilinear`` do the voxel work, the 2-D
masks (edge trim, S/N mask) stay small host arrays.  The semantics of the third-party calls are [RECALLED] and
restated in oracle/convolve.py.
"""
import numpy as np
import torch
from scipy.interpolate import CloughTocher2DInterpolator, griddata
from scipy.spatial import Delaunay

from . import signals
from .cube import SpecCube
from .engine import get_engine
from .utils import fits_io
from .utils.convolve import convolve_gauss2d
from .utils.fits_utils import downsample_header, get_pixel_mapping, linear_pixel_map

FWHM_TO_SIGMA = 1.0 / np.sqrt(8.0 * np.log(2.0))


class Beam(object):
    """Elliptical Gaussian beam (radio_beam.Beam stand-in): FWHM major / minor and position angle in degrees."""

    def __init__(self, major, minor=None, pa=None):
        self.major = float(major)
        self.minor = float(major if minor is None else minor)
        self.pa = 0.0 if pa is None else float(pa)
        if self.minor > self.major:
            raise ValueError("minor axis larger than the major axis")

    def deconvolve(self, other):
        """The beam that convolved with ``other`` gives this one (radio_beam ``Beam.deconvolve``, the AIPS / Miriad
        relations for two elliptical Gaussians at any position angles; convolve_tools.py:90-138 needs it when the
        target beam is not a scaled copy of the cube's beam) [RECALLED: radio_beam.utils.deconvolve].  For beams at
        the same angle the axes subtract in quadrature."""
        t1, t2 = np.radians(self.pa), np.radians(other.pa)
        a1, b1, a2, b2 = self.major ** 2, self.minor ** 2, other.major ** 2, other.minor ** 2
        alpha = a1 * np.cos(t1) ** 2 + b1 * np.sin(t1) ** 2 - a2 * np.cos(t2) ** 2 - b2 * np.sin(t2) ** 2
        beta = a1 * np.sin(t1) ** 2 + b1 * np.cos(t1) ** 2 - a2 * np.sin(t2) ** 2 - b2 * np.cos(t2) ** 2
        gamma = 2.0 * ((b1 - a1) * np.sin(t1) * np.cos(t1) - (b2 - a2) * np.sin(t2) * np.cos(t2))
        s = alpha + beta
        t = np.sqrt((alpha - beta) ** 2 + gamma ** 2)
        tol = 1e-12 * max(a1, a2)
        if alpha < -tol or beta < -tol or s - t <= tol:
            raise ValueError("beam could not be deconvolved")
        major, minor = np.sqrt(0.5 * (s + t)), np.sqrt(0.5 * (s - t))
        if abs(gamma) + abs(alpha - beta) <= tol:
            pa = 0.0 if major != minor else self.pa
        else:
            pa = np.degrees(0.5 * np.arctan2(-gamma, alpha - beta))
        if abs(((self.pa - other.pa) + 90.0) % 180.0 - 90.0) < 1e-9:
            pa = self.pa                                          # same angle: keep it exactly
        return Beam(major, minor, pa)

    def convolve(self, other):
        """The beam of this one convolved with ``other``: the second-moment matrices add."""
        def cov(b):
            t = np.radians(b.pa)
            c, s_ = np.cos(t), np.sin(t)
            R = np.array([[c, -s_], [s_, c]])
            return R @ np.diag([b.major ** 2, b.minor ** 2]) @ R.T
        w, vecs = np.linalg.eigh(cov(self) + cov(other))
        major, minor = np.sqrt(w[1]), np.sqrt(w[0])
        pa = np.degrees(np.arctan2(vecs[1, 1], vecs[0, 1])) if major != minor else 0.0
        return Beam(major, minor, ((pa + 90.0) % 180.0) - 90.0)

    def as_kernel(self, pixscale, support_scaling=8):
        """Sampled kernel in pixels: dict(kx, ky) [K] float32 for a circular beam, dict(k2) [K, K] otherwise
        (radio_beam EllipticalGaussian2DKernel: x_stddev along pa + 90 deg, support 8 sigma rounded up to odd)."""
        s_maj = self.major * FWHM_TO_SIGMA / pixscale
        s_min = self.minor * FWHM_TO_SIGMA / pixscale
        size = int(np.ceil(support_scaling * s_maj))
        size += 1 - size % 2
        r = np.arange(size, dtype=np.float64) - size // 2
        if s_maj == s_min:
            g = np.exp(-0.5 * (r / s_maj) ** 2)
            return dict(kx=g.astype(np.float32), ky=g.astype(np.float32))
        theta = np.radians(self.pa + 90.0)
        x, y = np.meshgrid(r, r)
        ct, st, s2t = np.cos(theta), np.sin(theta), np.sin(2 * theta)
        a = 0.5 * (ct ** 2 / s_maj ** 2 + st ** 2 / s_min ** 2)
        b = 0.5 * (s2t / s_maj ** 2 - s2t / s_min ** 2)
        c = 0.5 * (st ** 2 / s_maj ** 2 + ct ** 2 / s_min ** 2)
        return dict(k2=np.exp(-(a * x ** 2 + b * x * y + c * y ** 2)).astype(np.float32))


def _as_cube(cube):
    if isinstance(cube, SpecCube):
        return cube
    if isinstance(cube, str):
        return SpecCube.read(cube)
    raise TypeError("cube must be a SpecCube or a FITS path")


def _footprint(cube):
    """np.any(cube.mask.include(), axis=0) without the [nchan, ny, nx] temporary."""
    fp = np.zeros(cube.shape[1:], dtype=bool)
    for c0 in range(0, cube.shape[0], 64):
        fp |= np.isfinite(cube._data[c0:c0 + 64]).any(axis=0)
    return fp if cube.plane_include is None else fp & cube.plane_include


def edge_trim(cube, trim_width=3):
    """Mask the outer ``trim_width`` pixels of the cube's footprint (erosion with disk(trim_width))."""
    mask = _footprint(cube)
    if mask.size > 100:
        mask = signals.binary_erosion(mask, signals.disk(trim_width))
    return cube.with_mask(mask)


class _DeviceCube(object):
    """The cube as channel-major planes on the GPU (one upload), plus what the S/N mask needs from it."""

    def __init__(self, cube, device=None):
        self.eng = get_engine(device)
        host = cube._data if isinstance(cube, SpecCube) else cube
        self.planes = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)).to(self.eng.device, non_blocking=True)
        self.h2d_bytes = self.planes.numel() * 4

    @classmethod
    def wrap(cls, eng, planes):
        self = cls.__new__(cls)
        self.eng, self.planes, self.h2d_bytes = eng, planes, 0
        return self

    def mad_and_peak(self):
        """errmap = mad_std(cube, axis=0) and the peak per pixel, [ny, nx] float64 (b200fit_rms_snr)."""
        nchan, ny, nx = self.planes.shape
        rows = self.eng.gather(self.planes.view(nchan, ny * nx))
        out = self.eng.rms_snr(nchan, rows)
        packed = torch.stack([out["rms0"], out["peak"]]).cpu().numpy()
        return packed[0].reshape(ny, nx), packed[1].reshape(ny, nx)


def snr_mask(cube, snr_min=1.0, errmappath=None, _dev=None):
    """2-D mask of the pixels whose smoothed peak S/N exceeds ``snr_min``, cleaned by erosion / small-object removal
    and dilated by disk(3)."""
    dev = _dev if _dev is not None else _DeviceCube(_as_cube(cube))
    errmap, peak = dev.mad_and_peak()
    if errmappath is not None:
        errmap = fits_io.read(errmappath)[0]
    with np.errstate(all="ignore"):
        peaksnr = peak / errmap
    include = getattr(cube, 'plane_include', None)
    if include is not None:
        peaksnr[~include] = np.nan
    peaksnr = convolve_gauss2d(peaksnr, 1)
    with np.errstate(invalid="ignore"):
        planemask = peaksnr > snr_min
    if planemask.size > 100:
        planemask = signals.binary_erosion(planemask, signals.disk(1))
        planemask_im = signals.remove_small_objects(planemask, min_size=9)
        if np.sum(planemask_im) > 9:
            planemask = planemask_im
        planemask = signals.binary_dilation(planemask, signals.disk(3))
    return planemask


def convolve_sky(cube, beam, snrmasked=False, iterrefine=False, snr_min=3.0, device=None, _return_device=False):
    """The cube smoothed to ``beam`` (a Beam in the units of the header's BMAJ) on its own grid."""
    cube = _as_cube(cube)
    hdr = cube.header
    kern = beam.deconvolve(Beam(hdr['BMAJ'], hdr['BMIN'], hdr.get('BPA', 0.0))).as_kernel(abs(hdr['CDELT2']))
    ksize = next(iter(kern.values())).shape[0]
    if ksize > 65:
        raise ValueError("smoothing kernel of %d pixels exceeds the 65-pixel support of b200fit_convolve_planes "
                         "(beam %.3g pixels FWHM)" % (ksize, hdr['BMAJ'] / abs(hdr['CDELT2'])))
    dev = _DeviceCube(cube, device)
    eng = dev.eng
    kdev = {k: torch.from_numpy(v).to(eng.device) for k, v in kern.items()}
    mask = _footprint(cube)
    if snrmasked:
        planemask = snr_mask(cube, snr_min, _dev=dev)
        if np.sum(planemask) > 25:
            mask = mask & planemask

    def run(mask2d):
        inc = torch.from_numpy(np.ascontiguousarray(mask2d, dtype=np.uint8)).to(eng.device)
        return eng.convolve_planes(dev.planes, include=inc, **kdev)

    out = run(mask)
    if snrmasked and iterrefine:
        planemask = snr_mask(cube, snr_min, _dev=_DeviceCube.wrap(eng, out))
        if np.sum(planemask) > 25:
            mask2 = planemask if cube.plane_include is None else planemask & cube.plane_include
            out = run(mask2)
    if _return_device:
        return out, dev
    new = cube.with_data(out.cpu().numpy())
    new.header = dict(hdr, BMAJ=beam.major, BMIN=beam.minor, BPA=beam.pa)
    return new


def convolve_sky_byfactor(cube, factor, savename=None, edgetrim_width=5, downsample=True, device=None, **kwargs):
    """Smooth the cube to ``factor`` x its beam and (``downsample``) regrid it onto pixels ``factor`` x larger."""
    cube = _as_cube(cube)
    if edgetrim_width is not None:
        cube = edge_trim(cube, trim_width=edgetrim_width)
    hdr = cube.header
    if hdr.get('CUNIT1') != hdr.get('CUNIT2'):
        raise Exception("the spatial axis units for the cube do not match each other!")
    beam = Beam(hdr['BMAJ'] * factor, hdr['BMIN'] * factor, hdr.get('BPA', 0.0))
    out, dev = convolve_sky(cube, beam, device=device, _return_device=True, **kwargs)
    nhdr = dict(hdr, BMAJ=beam.major, BMIN=beam.minor, BPA=beam.pa)
    if downsample:
        nhdr = downsample_header(nhdr, factor=factor, axis=1)
        nhdr = downsample_header(nhdr, factor=factor, axis=2)
        nhdr['NAXIS1'] = int(np.rint(hdr['NAXIS1'] / factor))
        nhdr['NAXIS2'] = int(np.rint(hdr['NAXIS2'] / factor))
        (ay, by), (ax, bx) = linear_pixel_map(hdr, nhdr)
        out = dev.eng.resample_bilinear(out, nhdr['NAXIS2'], nhdr['NAXIS1'], ay, by, ax, bx)
    newcube = SpecCube(out.cpu().numpy(), cube.spectral_axis, rest_value=cube.rest_value, header=nhdr, unit=cube.unit)
    newcube.timing = dict(h2d_bytes=dev.h2d_bytes, d2h_bytes=out.numel() * 4)
    if savename is not None:
        newcube.write(savename, overwrite=True)
    return newcube


def get_celestial_hdr(header):
    """The on-sky cards of a header (WCS(header).celestial.to_header() + NAXIS1/2)."""
    keep = ("NAXIS1", "NAXIS2", "CTYPE1", "CTYPE2", "CRVAL1", "CRVAL2", "CRPIX1", "CRPIX2", "CDELT1", "CDELT2",
            "CUNIT1", "CUNIT2", "CROTA2", "PC1_1", "PC1_2", "PC2_1", "PC2_2", "CD1_1", "CD1_2", "CD2_1", "CD2_2",
            "RADESYS", "EQUINOX", "LONPOLE", "LATPOLE")
    return {k: header[k] for k in keep if k in header}


def regrid(image, header1, header2, dmask=None, method='cubic'):
    """``image`` (on header1) interpolated onto the grid of header2 with scipy griddata over its finite pixels."""
    grid1 = get_pixel_mapping(header1, header2)
    X, Y = np.meshgrid(np.arange(image.shape[1]), np.arange(image.shape[0]))
    mask = np.isfinite(image)
    if dmask is None:
        dmask = np.ones(grid1[0].shape, dtype=bool)
    return griddata((X[mask], Y[mask]), image[mask], (grid1[1] * dmask, grid1[0] * dmask), method=method,
                    fill_value=np.nan)


def regrid_many(images, header1, header2, method='cubic'):
    """``regrid`` of several maps [k, ny, nx] at once, with results identical to k separate calls: maps that share
    their finite footprint share one Delaunay triangulation and one point location (the triangulation is 60 % of a
    cubic griddata call on a 256 x 256 map, the point location most of the rest)."""
    images = np.asarray(images)
    if method != 'cubic':
        return np.array([regrid(im, header1, header2, method=method) for im in images])
    grid1 = get_pixel_mapping(header1, header2)
    X, Y = np.meshgrid(np.arange(images.shape[2]), np.arange(images.shape[1]))
    out = np.empty((images.shape[0],) + grid1[0].shape)
    finite = np.isfinite(images)
    done = np.zeros(images.shape[0], bool)
    for k in range(images.shape[0]):
        if done[k]:
            continue
        same = [j for j in range(k, images.shape[0]) if not done[j] and np.array_equal(finite[j], finite[k])]
        m = finite[k]
        tri = Delaunay(np.column_stack([X[m], Y[m]]))
        ip = CloughTocher2DInterpolator(tri, np.stack([images[j][m] for j in same], axis=1), fill_value=np.nan)
        res = ip(grid1[1], grid1[0])
        for i, j in enumerate(same):
            out[j] = res[..., i]
            done[j] = True
    return out
