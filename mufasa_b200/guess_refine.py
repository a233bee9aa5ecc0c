"""Guess refinement on the host: mirror of the piece of mufasa/guess_refine.py the fit driver calls
(tautex_renorm :183-213, from cubefit_gen multi_v_fit.py:330-331)."""
import numpy as np

from . import moment_guess as mmg

nu0_nh3 = 23.6944955


def tautex_renorm(taumap, texmap, tau_thresh=0.21, tex_thresh=15.0, nu=nu0_nh3):
    """Re-normalise the degenerate (tau, Tex) pair of optically thin guesses: tau below ``tau_thresh`` is raised to
    the threshold and Tex lowered so the peak brightness is kept.  In place; returns (taumap, texmap).

    Reference quirks kept: get_tex is called WITHOUT nu (its default 23.722634 GHz applies, guess_refine.py:199), and
    the reference's second, high-Tex branch assigns through chained fancy indexing (:209-210) and so changes
    nothing -- it is omitted here."""
    f_tau = 0.5
    isthin = np.logical_and(taumap < tau_thresh, np.isfinite(taumap))
    TA_lowtau = mmg.peakT(texmap[isthin], taumap[isthin] * f_tau, nu=nu)
    texmap[isthin] = mmg.get_tex(TA_lowtau, tau=tau_thresh * f_tau)
    taumap[isthin] = tau_thresh
    return taumap, texmap
