"""SpecCube: the minimal stand-in for ``spectral_cube.SpectralCube`` the fit path needs.

The reference hands a SpectralCube in K / km/s to ``UltraCube`` (UltraCube.py:143-171, ``to_K`` :1850-1932);
astropy / spectral_cube are absent here and are not on the hot path, so the drivers work on this plain
container: ``_data[nchan, ny, nx]`` (float32 as FITS stores it), ``spectral_axis`` in km/s (uniform), a header
dict, and the bad-value mask ``isfinite`` (spectral_cube's default LazyMask).
"""
import numpy as np

from .pcube import VelocityAxis


class _Mask(object):
    def __init__(self, data, plane=None):
        self._data = data
        self._plane = plane

    def include(self):
        ok = np.isfinite(self._data)
        return ok if self._plane is None else ok & self._plane[None]


class SpecCube(object):
    def __init__(self, data, v_kms, rest_value=None, header=None, unit='K'):
        data = np.asarray(data)
        if data.ndim != 3:
            raise ValueError("cube must be [nchan, ny, nx]")
        if data.dtype != np.float32:
            data = data.astype(np.float32)
        self._data = data
        self.spectral_axis = np.asarray(v_kms, dtype=np.float64)
        if self.spectral_axis.shape != (data.shape[0],):
            raise ValueError("spectral axis length does not match the cube")
        self.rest_value = rest_value          # Hz, None => the model's rest frequency (multi_v_fit.py:109-113)
        self.header = dict(header or {})
        self.unit = unit
        self.plane_include = None             # [ny, nx] bool set by with_mask: excluded pixels read as NaN
        self.mask = _Mask(data)

    @property
    def shape(self):
        return self._data.shape

    @property
    def filled_data(self):
        return self._data

    def xarr(self):
        return VelocityAxis(self.spectral_axis, refX=self.rest_value)

    def with_mask(self, plane_mask):
        """A view of the same voxels with the pixels outside ``plane_mask`` [ny, nx] masked (spectral_cube's
        ``with_mask`` for the 2-D masks convolve_tools builds; no copy of the data)."""
        new = SpecCube.__new__(SpecCube)
        new.__dict__.update(self.__dict__)
        plane_mask = np.asarray(plane_mask, bool)
        if plane_mask.shape != self._data.shape[1:]:
            raise ValueError("with_mask takes a [ny, nx] mask")
        new.plane_include = plane_mask if self.plane_include is None else (plane_mask & self.plane_include)
        new.mask = _Mask(self._data, new.plane_include)
        return new

    @classmethod
    def read(cls, path):
        """Load a FITS cube (primary HDU, [nchan, ny, nx]) whose third axis is a radio velocity (VRAD / VELO*, m/s
        or km/s) or a frequency (FREQ, Hz; converted with the radio convention about RESTFRQ, as
        multi_v_fit.register_pcube does, multi_v_fit.py:109-116)."""
        from .utils import fits_io
        data, hdr = fits_io.read(path)
        header = {k: v[0] for k, v in hdr.items()}
        if data.ndim == 4 and data.shape[0] == 1:
            data = data[0]
        n = data.shape[0]
        axis = header.get('CRVAL3', 0.0) + (np.arange(n) + 1 - header.get('CRPIX3', 1.0)) * header.get('CDELT3', 1.0)
        ctype = str(header.get('CTYPE3', 'VRAD')).upper()
        cunit = str(header.get('CUNIT3', 'Hz' if ctype.startswith('FREQ') else 'm/s')).strip().lower()
        rest = header.get('RESTFRQ', header.get('RESTFREQ'))
        if ctype.startswith('FREQ'):
            if rest is None:
                raise ValueError("frequency axis without RESTFRQ")
            scale = {'hz': 1.0, 'khz': 1e3, 'mhz': 1e6, 'ghz': 1e9}[cunit]
            v_kms = 299792.458 * (1.0 - axis * scale / float(rest))
        elif ctype.startswith(('VRAD', 'VELO')):
            v_kms = axis * {'m/s': 1e-3, 'm s-1': 1e-3, 'km/s': 1.0, 'km s-1': 1.0}[cunit]
        else:
            raise ValueError("unsupported spectral axis %r" % ctype)
        return cls(data, v_kms, rest_value=None if rest is None else float(rest), header=header,
                   unit=str(header.get('BUNIT', 'K')).strip() or 'K')

    def write(self, path, overwrite=True):
        import os
        from .utils import fits_io
        if not overwrite and os.path.exists(path):
            raise IOError("%s exists" % path)
        hdr = dict(self.header)
        v = self.spectral_axis
        hdr.update(CTYPE3='VRAD', CUNIT3='km/s', CRPIX3=1.0, CRVAL3=float(v[0]),
                   CDELT3=float((v[-1] - v[0]) / max(len(v) - 1, 1)), BUNIT=self.unit)
        if self.rest_value is not None:
            hdr['RESTFRQ'] = float(self.rest_value)
        fits_io.write(path, self._data, hdr)

    def with_data(self, data):
        return SpecCube(data, self.spectral_axis, rest_value=self.rest_value, header=self.header, unit=self.unit)
