"""Frequency-space hyperfine radiative-transfer model (oracle).  TEST INFRASTRUCTURE.

Restates, in the reference's own operation order (fp64, numpy, Python loop over hyperfines):
  * HyperfineModel.multi_v_spectrum      mufasa/spec_models/HyperfineModel.py:22-58
  * HyperfineModel._single_spectrum_hf   mufasa/spec_models/HyperfineModel.py:60-109
  * BaseModel.T_antenna                  mufasa/spec_models/BaseModel.py:176-193
The spectral axis is what pyspeckit hands the model: frequency in GHz, radio convention
nu_i = nu_ref * (1 - V_i / c)  (multi_v_fit.py:109-116 sets refX/velocity_convention).
"""
import numpy as np

from . import constants as K


def vaxis(nchan, dv, v0=None):
    """Uniform velocity axis of the synthetic cubes: V_i = (i - nchan/2) * dv  (SURVEY App. E)."""
    if v0 is None:
        v0 = -(nchan // 2) * dv
    return v0 + dv * np.arange(nchan, dtype=float)


def freq_axis_ghz(v_kms, nu_ref_hz):
    """Radio-convention velocity -> frequency (GHz)."""
    return nu_ref_hz * (1.0 - np.asarray(v_kms, dtype=float) / K.CKMS) / 1e9


def T_antenna(Tbright, nu_ghz):
    """BaseModel.py:176-193."""
    T0 = (K.H_CGS * nu_ghz * 1e9 / K.KB_CGS)
    return T0 / (np.exp(T0 / Tbright) - 1)


def single_spectrum_hf(xghz, tex, tau, width, xoff_v, nu0_hz, voff_lines, tau_wts, background_ta=0.0):
    """HyperfineModel.py:60-109 for one line name (line_names has exactly one entry, BaseModel.py:59-61)."""
    runspec = np.zeros(len(xghz))
    voff_lines = np.array(voff_lines, dtype=float)
    tau_wts = np.array(tau_wts, dtype=float)
    lines = (1 - voff_lines / K.CKMS) * nu0_hz / 1e9
    tau_wts = tau_wts / tau_wts.sum()
    nuwidth = np.abs(width / K.CKMS * lines)
    nuoff = xoff_v / K.CKMS * lines

    tauprof = np.zeros(len(xghz))
    for kk, nuo in enumerate(nuoff):
        tauprof += (tau * tau_wts[kk] *
                    np.exp(-(xghz + nuo - lines[kk]) ** 2 /
                           (2.0 * nuwidth[kk] ** 2)))

    T0 = (K.H_CGS * xghz * 1e9 / K.KB_CGS)
    runspec += ((T0 / (np.exp(T0 / tex) - 1) * (1 - np.exp(-tauprof)) +
                 background_ta * np.exp(-tauprof)))
    return runspec


def multi_v_spectrum(xghz, pars, fittype):
    """HyperfineModel.py:22-58: chain components, component 1 is the rear slab."""
    nu0, voff, wts = K.line(fittype)
    xghz = np.asarray(xghz, dtype=float)
    pars = np.asarray(pars, dtype=float)
    background_ta = T_antenna(K.TCMB, xghz)
    model_spectrum = background_ta
    for vel, width, tex, tau in zip(pars[::4], pars[1::4], pars[2::4], pars[3::4]):
        model_spectrum = single_spectrum_hf(xghz, tex, tau, width, vel, nu0, voff, wts,
                                            background_ta=background_ta)
        background_ta = model_spectrum
    return model_spectrum - T_antenna(K.TCMB, xghz)


def model_on_vaxis(v_kms, pars, fittype, nu_ref_hz=None):
    """Convenience: evaluate on a velocity axis (radio convention about nu_ref, default = model nu0)."""
    nu0, _, _ = K.line(fittype)
    return multi_v_spectrum(freq_axis_ghz(v_kms, nu0 if nu_ref_hz is None else nu_ref_hz), pars, fittype)


def modelcube(v_kms, parcube, fittype, nu_ref_hz=None, dtype=np.float64):
    """pyspeckit Cube.get_modelcube [RECALLED]: model for every pixel with finite, not-all-zero
    parameters; NaN elsewhere; array dtype follows the data cube (``np.full_like(cube, nan)``)."""
    parcube = np.asarray(parcube, dtype=float)
    npar, ny, nx = parcube.shape
    out = np.full((len(v_kms), ny, nx), np.nan, dtype=dtype)
    nu0, _, _ = K.line(fittype)
    xghz = freq_axis_ghz(v_kms, nu0 if nu_ref_hz is None else nu_ref_hz)
    valid = np.any(parcube != 0, axis=0) & np.all(np.isfinite(parcube), axis=0)
    for y, x in zip(*np.nonzero(valid)):
        out[:, y, x] = multi_v_spectrum(xghz, parcube[:, y, x], fittype)
    return out


# ---------------------------------------------------------------------------------------------
# Velocity-space restatement + analytic Jacobian.  Used by tests to check the
# formulas the CUDA kernel implements; NOT what the reference evaluates.
def model_and_jacobian_vspace(v_kms, pars, fittype):
    nu0, voff, wts = K.line(fittype)
    w = wts / wts.sum()
    V = np.asarray(v_kms, dtype=float)
    pars = np.asarray(pars, dtype=float)
    ncomp = len(pars) // 4
    nu = nu0 * (1.0 - V / K.CKMS)
    T0 = K.H_CGS * nu / K.KB_CGS
    Jbg = T0 / np.expm1(T0 / K.TCMB)
    S = Jbg.copy()
    dS = np.zeros((4 * ncomp, len(V)))
    a = 1.0 - voff / K.CKMS
    for c in range(ncomp):
        v, sig, tex, tau0 = pars[4 * c:4 * c + 4]
        u = V[:, None] - (v + voff[None, :] * (1.0 - v / K.CKMS))
        s = sig * a[None, :]
        g = w[None, :] * np.exp(-u ** 2 / (2 * s ** 2))
        G = g.sum(1)
        dtau_dv = tau0 * (g * u / s ** 2 * a[None, :]).sum(1)
        dtau_ds = tau0 * (g * u ** 2 / (s ** 2 * sig)).sum(1)
        tau = tau0 * G
        e = np.exp(-tau)
        ex = np.exp(T0 / tex)
        Jt = T0 / (ex - 1.0)
        dJ = (T0 / tex) ** 2 * ex / (ex - 1.0) ** 2
        D = (Jt - S) * e
        dS[:4 * c] *= e[None, :]
        dS[4 * c + 0] = D * dtau_dv
        dS[4 * c + 1] = D * dtau_ds
        dS[4 * c + 2] = dJ * (1.0 - e)
        dS[4 * c + 3] = D * G
        S = Jt * (1.0 - e) + S * e
    return S - Jbg, dS
