"""Moment-based initial guesses: host mirror of mufasa/moment_guess.py (moment_guesses :388-518, LineSetup :27-58,
get_tex / get_tau / peakT :774-800).  Produces the per-pixel guesses the CUDA fit kernel stages."""
import numpy as np

# moment_guess.py:15 imports h, kb from pyspeckit's ammonia_constants; CODATA-2018 cgs values
h = 6.62607015e-27
kb = 1.380649e-16
TBG = 2.7315


class LineSetup(object):
    """Guessing ranges of Tex and tau per species (moment_guess.py:27-58)."""

    def __init__(self, linetype='nh3', new_recipe=True):
        if linetype not in ('nh3', 'n2hp'):
            raise Exception("{} is an invalid linetype".format(linetype))
        if new_recipe:
            self.tex_max, self.tau_max, self.tex_min, self.tau_min = 8.0, 5.0, 3.0, 0.1
        else:
            self.tex_max, self.tau_max, self.tex_min, self.tau_min = 8.0, 1.0, 3.1, 0.3


def moment_guesses(moment1, moment2, ncomp, sigmin=0.07, tex_guess=3.2, tau_guess=0.5, moment0=None, linetype='nh3',
                   mom0_floor=None):
    """[v, sigma, Tex, tau] x ncomp from moment maps.  Same recipe and the same quirks as the reference: ``moment2``
    is used as a width whatever it is, gs_sig clamps IN PLACE into moment2/ncomp, and the Tex/tau range clamps happen
    after the values were copied into the output (they do not affect it)."""
    ls = LineSetup(linetype)
    if moment0 is not None:
        norm_ref = 99.73
        mom0high = np.percentile(moment0[np.isfinite(moment0)], norm_ref)
        if mom0_floor is None:
            m0n = moment0.copy() * norm_ref / 100.0 / mom0high
            tex_guess = m0n * ls.tex_max
            tau_guess = m0n * ls.tau_max
        else:
            m0n = moment0.copy() * norm_ref / 100.0 / (mom0high - mom0_floor)
            tex_guess = m0n * (ls.tex_max - ls.tex_min) + ls.tex_min
            tau_guess = m0n * (ls.tau_max - ls.tau_min) + ls.tau_min
    m1, m2 = moment1, moment2
    gs_sig = m2 / ncomp
    with np.errstate(invalid="ignore"):
        gs_sig[gs_sig < sigmin] = sigmin
    gg = np.zeros((ncomp * 4,) + m1.shape)
    if ncomp == 1:
        gg[0], gg[1], gg[2], gg[3] = m1, gs_sig, tex_guess, tau_guess
    elif ncomp == 2:
        sigmaoff, tau2_frac = 0.4, 0.25
        gg[0] = m1 - sigmaoff * m2
        gg[1] = gs_sig
        gg[2] = tex_guess
        gg[3] = tau_guess * (1 - tau2_frac)
        gg[4] = m1 + sigmaoff * m2
        gg[5] = gs_sig
        gg[6] = tex_guess * 0.8
        gg[7] = tau_guess * tau2_frac
    else:
        raise NotImplementedError("the CUDA fitter covers 1- and 2-component models")
    return gg


def get_tex(Ta, tau=0.5, nu=23.722634):
    T0 = (h * nu * 1e9 / kb)
    term1 = Ta / T0 / (1 - np.exp(-tau))
    term2 = 1.0 / (np.exp(T0 / TBG) - 1)
    return T0 / np.log(1.0 / (term1 + term2) + 1)


def get_tau(Ta, tex=6.0, nu=23.722634):
    T0 = (h * nu * 1e9 / kb)
    term1 = 1 / (np.exp(T0 / tex) - 1) - 1 / (np.exp(T0 / TBG) - 1)
    return -np.log(-Ta / T0 / term1 + 1)


def peakT(tex, tau, nu=23.722634):
    T0 = (h * nu * 1e9 / kb)
    return (T0 / (np.exp(T0 / tex) - 1) - T0 / (np.exp(T0 / TBG) - 1)) * (1 - np.exp(-tau))
