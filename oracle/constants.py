"""Constants of the hot path (oracle side).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Sources:
  * TCMB, h, kb, c: mufasa/spec_models/BaseModel.py:13-15,25-29 (astropy CODATA-2018 cgs values).
  * N2H+ (1-0): mufasa/spec_models/m_constants.py:22-61 (values restated from that table).
  * NH3 (1,1): imported by the reference from pyspeckit (m_constants.py:11-17), which is NOT in the
    reference tree -> values restated from pyspeckit's published ammonia_constants ('oneone').
    [RECALLED -- parity unpinned for this table]
  * limits: mufasa/spec_models/meta_model.py:95-101,128-132.
"""
import numpy as np

TCMB = 2.7315                 # K        BaseModel.py:13
H_CGS = 6.62607015e-27        # erg s    BaseModel.py:14 (astropy constants.h.cgs)
KB_CGS = 1.380649e-16         # erg / K  BaseModel.py:15 (astropy constants.k_B.cgs)
CKMS = 299792.458             # km / s   BaseModel.py:26

LINES = {
    # fittype -> dict(nu0 [Hz], voff [km/s], wts)
    "nh3_multi_v": dict(
        linename="oneone",
        nu0=23.6944955e9,
        voff=[19.8513, 19.3159, 7.88669, 7.46967, 7.35132, 0.460409, 0.322042,
              -0.0751680, -0.213003, 0.311034, 0.192266, -0.132382, -0.250923,
              -7.23349, -7.37280, -7.81526, -19.4117, -19.5500],
        wts=[0.0740740, 0.148148, 0.0925930, 0.166667, 0.0185190, 0.0370370,
             0.0185190, 0.0185190, 0.0925930, 0.0333330, 0.300000, 0.466667,
             0.0333330, 0.0925930, 0.0185190, 0.166667, 0.0740740, 0.148148],
    ),
    "n2hp_multi_v": dict(
        linename="onezero",
        nu0=93173.7637e6,
        voff=[-7.9930, -7.9930, -7.9930, -0.6112, -0.6112, -0.6112, 0.0000, 0.9533,
              0.9533, 5.5371, 5.5371, 5.5371, 5.9704, 5.9704, 6.9238],
        wts=[0.025957, 0.065372, 0.019779, 0.004376, 0.034890, 0.071844, 0.259259,
             0.156480, 0.028705, 0.041361, 0.013309, 0.056442, 0.156482, 0.028705,
             0.037038],
    ),
}
# deprecated module names the reference still accepts as shims (ammonia_multiv.py:24-30, n2hp_multiv.py:23-29)
ALIASES = {"ammonia_multiv": "nh3_multi_v", "n2hp_multiv": "n2hp_multi_v"}


def limits(fittype):
    """(Texmin, Texmax, sigmin, sigmax, taumin, taumax, eps)  meta_model.py:95-101,128."""
    fittype = ALIASES.get(fittype, fittype)
    taumax = 40.0 if fittype == "n2hp_multi_v" else 30.0
    return dict(Texmin=3.0, Texmax=40.0, sigmin=0.07, sigmax=2.5, taumin=0.1, taumax=taumax, eps=0.001)


def line(fittype):
    fittype = ALIASES.get(fittype, fittype)
    d = LINES[fittype]
    return d["nu0"], np.asarray(d["voff"], dtype=float), np.asarray(d["wts"], dtype=float)
