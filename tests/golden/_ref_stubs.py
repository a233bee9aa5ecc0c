"""Import the reference's own Python sources (from /root/reference) in a container that lacks its
third-party imports.  Used ONLY by tests/golden/make_golden.py to generate committed golden vectors.

The reference's arithmetic on the hot path is plain numpy/scipy; astropy / pyspeckit / spectral_cube / dask /
skimage / ... are needed only for units, I/O, the solver driver and plotting.  This module installs
permissive stand-in modules for those names so that the reference *files themselves* can be executed:
  * ``astropy.constants`` h, k_B, c  -> CODATA-2018 values (what astropy ships)
  * ``astropy.units``                -> a 20-line unit/Quantity sufficient for ``x * u.Hz`` and ``.to('GHz').value``
  * ``pyspeckit.spectrum.models.ammonia{,_constants}`` -> the NH3 (1,1) table  [RECALLED from pyspeckit;
    injected from oracle/constants.py -- this is the one input that is not the reference's own]
  * ``dask.array`` -> a module whose ``Array`` class nothing is an instance of (so the numpy branches run)
  * everything else -> inert dummies
Nothing here is shipped or imported by the product or by the test-suite proper.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np

STUB_ROOTS = ("astropy", "pyspeckit", "spectral_cube", "dask", "skimage", "radio_beam", "reproject",
              "matplotlib", "plotly")


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_dummy(name)

    def _binop(cls, other):
        return cls
    __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __add__ = __radd__ = __sub__ = __rsub__ = __pow__ = _binop


def _make_dummy(name):
    return _DummyMeta(name, (object,), {
        "__init__": lambda self, *a, **k: None,
        "__call__": lambda self, *a, **k: self,
        "__getattr__": lambda self, n: _make_dummy(n)(),
        "__getitem__": lambda self, k: self,
        "__iter__": lambda self: iter(()),
    })


class Unit(object):
    _to_hz = {"Hz": 1.0, "kHz": 1e3, "MHz": 1e6, "GHz": 1e9}

    def __init__(self, name):
        self.name = name

    def __truediv__(self, other):
        return Unit("%s / %s" % (self.name, other.name))

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit("%s %s" % (self.name, other.name))
        return Quantity(other, self)

    __rmul__ = __mul__

    def to_string(self):
        return self.name

    def __eq__(self, other):
        return getattr(other, "name", other) == self.name

    def __hash__(self):
        return hash(self.name)


class Quantity(object):
    def __init__(self, value, unit):
        self.value = value
        self.unit = unit

    def to(self, unit):
        name = unit if isinstance(unit, str) else unit.name
        if name == self.unit.name:
            return Quantity(self.value, Unit(name))
        f = Unit._to_hz[self.unit.name] / Unit._to_hz[name]
        return Quantity(self.value * f, Unit(name))


class _Const(object):
    def __init__(self, cgs, si_kms=None):
        self.cgs = types.SimpleNamespace(value=cgs)
        self._kms = si_kms

    def to(self, unit):
        if unit.name == "km / s":
            return Quantity(self._kms, unit)
        if unit.name == "cm / s":
            return Quantity(self._kms * 1e5, unit)
        raise ValueError(unit.name)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        d = _make_dummy(name)
        setattr(self, name, d)
        return d


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


def _populate(module):
    name = module.__name__
    if name == "astropy.units":
        for u in ("Hz", "kHz", "MHz", "GHz", "km", "cm", "s", "K", "m", "deg", "arcsec", "pix", "dimensionless_unscaled"):
            setattr(module, u, Unit(u))
        module.UnitConversionError = type("UnitConversionError", (Exception,), {})
        module.UnitScaleError = type("UnitScaleError", (Exception,), {})
        module.Quantity = Quantity
    elif name == "astropy.constants":
        module.h = _Const(6.62607015e-27)           # CODATA 2018, erg s
        module.k_B = _Const(1.380649e-16)           # erg / K
        module.c = _Const(2.99792458e10, si_kms=299792.458)
    elif name == "astropy":
        import importlib
        module.units = importlib.import_module("astropy.units")
        module.constants = importlib.import_module("astropy.constants")
    elif name in ("pyspeckit.spectrum.models.ammonia", "pyspeckit.spectrum.models.ammonia_constants"):
        from oracle import constants as K            # the RECALLED NH3 table (see module docstring)
        d = K.LINES["nh3_multi_v"]
        module.line_names = ["oneone"]
        module.freq_dict = {"oneone": d["nu0"]}
        module.voff_lines_dict = {"oneone": list(d["voff"])}
        module.tau_wts_dict = {"oneone": list(d["wts"])}
        module.ckms = K.CKMS
        module.ccms = K.CKMS * 1e5
        module.h = K.H_CGS
        module.kb = K.KB_CGS
    elif name == "spectral_cube.utils":
        module.SpectralCubeWarning = type("SpectralCubeWarning", (Warning,), {})
        module.NoBeamError = type("NoBeamError", (Exception,), {})
    elif name == "dask.array":
        module.Array = type("Array", (object,), {})
    elif name == "dask.array.core":
        module.Array = sys.modules["dask.array"].Array if "dask.array" in sys.modules else type("Array", (object,), {})


def install(reference_root="/root/reference"):
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


class FakeXarr(object):
    """Stands in for pyspeckit's SpectroscopicAxis: HyperfineModel.py:41-42 only asks for unit/value/len."""

    def __init__(self, ghz):
        self.value = np.asarray(ghz, dtype=float)
        self.unit = Unit("GHz")

    def __len__(self):
        return len(self.value)

    def as_unit(self, unit):
        return self


class FakeCube(object):
    """Stands in for SpectralCube in UltraCube.get_rss/get_residual (:1384-1475, :1743-1800):
    only ``filled_data[:].value``, ``_data`` and ``shape`` are touched."""

    def __init__(self, data):
        self._data = data
        self.shape = data.shape

    @property
    def filled_data(self):
        data = self._data

        class _F(object):
            def __getitem__(self, key):
                return types.SimpleNamespace(value=data[key])
        return _F()
