"""Per-pixel fit driver (oracle).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates what the reference's ``pcube.fiteach(...)`` call does to every pixel
(call sites multi_v_fit.py:374-393, :192-206, retry :456-478); the callee is pyspeckit
``Cube.fiteach`` -> ``Specfit`` -> ``SpectralModel.fitter`` -> ``mpfit``  [RECALLED, parity unpinned]:

  * valid pixel = maskmap & (any finite channel) & (all guesses finite)
  * weights: MUFASA passes neither errmap nor errorcube (multi_v_fit.py:234-238 loads errmap but never
    forwards it) => err_i = std(spectrum[finite]) (population std, in the cube's dtype), same for every channel
  * non-finite channels: data -> 0, err -> 1e10 (no weight)
  * residual f(p) = (y - model(x; p)) / err on the frequency axis; mpfit defaults (see oracle/mpfit.py)
  * success: parcube = mp.params, errcube = mp.perror, has_fit = True (also for status 5 = maxiter);
    mp.status <= 0 or any exception: zeros and has_fit = False
  * pixel order (start_from_point) is irrelevant because use_neighbor_as_guess=False (multi_v_fit.py:381)
"""
import multiprocessing as mp
import os

import numpy as np

from . import constants as K
from . import model as M
from . import mpfit as MP


def fit_spectrum(y, xghz, guess, minpars, maxpars, fittype, maxiter=200):
    """Fit one spectrum.  Returns (params[P], perror[P], chi2, status, nfev, err_used)."""
    y = np.asarray(y)
    npar = len(guess)
    finite = y == y
    zero = (np.zeros(npar), np.zeros(npar), 0.0, 0, 0, np.nan)
    if not np.any(finite) or not np.all(np.isfinite(guess)):
        return zero
    err_val = y[finite].std()                       # dtype of the cube, as numpy would compute it
    if not np.isfinite(err_val) or err_val <= 0:
        return zero
    data = np.where(finite, y, 0).astype(np.float64)
    err = np.where(finite, np.float64(err_val), 1e10)

    def deviates(p):
        return (data - M.multi_v_spectrum(xghz, p, fittype)) / err

    limited = np.ones((npar, 2), dtype=bool)
    limits = np.stack([np.asarray(minpars, float), np.asarray(maxpars, float)], axis=1)
    try:
        with np.errstate(all="ignore"):
            res = MP.mpfit(deviates, np.asarray(guess, float), limited, limits, maxiter=maxiter)
    except Exception:                                # pyspeckit swallows per-pixel failures
        return zero
    if res.status <= 0 or res.perror is None:
        return (np.zeros(npar), np.zeros(npar), 0.0, int(res.status), res.nfev, float(err_val))
    # chi2 AT THE RETURNED PARAMETERS.  mpfit's own ``fnorm`` is max(fnorm(final), fnorm1(last trial))**2, i.e. the
    # chi2 of a REJECTED step whenever the fit ends on one; MUFASA never reads it (pyspeckit stores no per-pixel
    # chi2; UltraCube.get_chisq / get_rss re-evaluate the model at parcube), so parity is defined on this value.
    chi2 = float(np.sum(deviates(res.params) ** 2))
    return res.params, res.perror, chi2, int(res.status), res.nfev, float(err_val)


_G = {}


def _work(idx):
    cube, xghz, guesses, lo, hi, fittype, maxiter = _G["args"]
    out = []
    for (y, x) in idx:
        out.append(fit_spectrum(cube[:, y, x], xghz, guesses[:, y, x], lo, hi, fittype, maxiter))
    return out


def fiteach(cube, v_kms, guesses, minpars, maxpars, fittype, maskmap=None, multicore=1,
            nu_ref_hz=None, maxiter=200):
    """Returns dict(parcube[P,ny,nx], errcube[P,ny,nx], has_fit[ny,nx], chi2, status, nfev, err)."""
    cube = np.asarray(cube)
    nchan, ny, nx = cube.shape
    guesses = np.asarray(guesses, dtype=float)
    npar = guesses.shape[0]
    nu0, _, _ = K.line(fittype)
    xghz = M.freq_axis_ghz(v_kms, nu0 if nu_ref_hz is None else nu_ref_hz)
    ok = np.any(np.isfinite(cube), axis=0) & np.all(np.isfinite(guesses), axis=0)
    if maskmap is not None:
        ok &= np.asarray(maskmap, dtype=bool)
    pix = list(zip(*np.nonzero(ok)))
    out = dict(parcube=np.zeros((npar, ny, nx)), errcube=np.zeros((npar, ny, nx)),
               has_fit=np.zeros((ny, nx), dtype=bool), chi2=np.zeros((ny, nx)),
               status=np.zeros((ny, nx), dtype=np.int32), nfev=np.zeros((ny, nx), dtype=np.int32),
               err=np.full((ny, nx), np.nan))
    if not pix:
        return out
    _G["args"] = (cube, xghz, guesses, np.asarray(minpars, float), np.asarray(maxpars, float), fittype, maxiter)
    multicore = max(1, int(multicore))
    if multicore == 1 or len(pix) < 2 * multicore:
        results = _work(pix)
    else:
        nchunk = min(len(pix), multicore * 8)
        chunks = [pix[i::nchunk] for i in range(nchunk)]
        ctx = mp.get_context("fork")                 # workers inherit _G (pyspeckit parallel_map also forks)
        with ctx.Pool(multicore) as pool:
            parts = pool.map(_work, chunks)
        results = [None] * len(pix)
        for i, part in enumerate(parts):
            results[i::nchunk] = part
    for (y, x), (p, e, c2, st, nf, er) in zip(pix, results):
        out["parcube"][:, y, x] = p
        out["errcube"][:, y, x] = e
        out["chi2"][y, x] = c2
        out["status"][y, x] = st
        out["nfev"][y, x] = nf
        out["err"][y, x] = er
        out["has_fit"][y, x] = st > 0
    return out


def default_cores():
    """Reference default: cpu_count() - 1 workers (mufasa/utils/multicore.py:5-18)."""
    return max(1, (os.cpu_count() or 2) - 1)
