"""SpecCube: the minimal stand-in for ``spectral_cube.SpectralCube`` the fit path needs.

The reference hands a SpectralCube in K / km/s to ``UltraCube`` (UltraCube.py:143-171, ``to_K`` :1850-1932);
astropy / spectral_cube are absent here and are not on the hot path, so the drivers work on this plain
container: ``_data[nchan, ny, nx]`` (float32 as FITS stores it), ``spectral_axis`` in km/s (uniform), a header
dict, and the bad-value mask ``isfinite`` (spectral_cube's default LazyMask).
"""
import numpy as np

from .pcube import VelocityAxis


class _Mask(object):
    def __init__(self, data):
        self._data = data

    def include(self):
        return np.isfinite(self._data)


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
        self.mask = _Mask(data)

    @property
    def shape(self):
        return self._data.shape

    @property
    def filled_data(self):
        return self._data

    def xarr(self):
        return VelocityAxis(self.spectral_axis, refX=self.rest_value)

    def with_data(self, data):
        return SpecCube(data, self.spectral_axis, rest_value=self.rest_value, header=self.header, unit=self.unit)
