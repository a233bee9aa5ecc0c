"""Shared helpers of the parity tests: sampled tiles of BASELINE.json's synthetic configs, oracle fits,
and the three north-star criteria."""
import ctypes as C
import os
import subprocess

import numpy as np

from mufasa_b200 import _abi, synth
from oracle import constants as K
from oracle import fiteach as F
from oracle import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_modelcube(parcube, v, fittype):
    return M.modelcube(v, parcube, fittype)


def sample_tile(cfg, nrows=4, xstep=8, row_lo=20):
    """Rows spread over config ``cfg``, every ``xstep``-th column: covers truth ncomp 0/1/2 regions.
    Returns dict(cube[nchan, nrows, nxs] f32, v, fittype, dv, guess1, guess2, ncomp_true)."""
    fittype, NY, NX, nchan, dv = synth.CONFIGS[cfg]
    rows = np.linspace(row_lo, NY - 1 - row_lo, nrows).astype(int)
    parts = [synth.make_cube(cfg, oracle_modelcube, rows=(int(y), int(y) + 1)) for y in rows]
    out = dict(v=parts[0]["v"], fittype=fittype, dv=dv, nchan=nchan)
    for k in ("cube", "guess1", "guess2", "truth2"):
        out[k] = np.ascontiguousarray(np.concatenate([p[k][:, :, ::xstep] for p in parts], axis=1))
    out["ncomp_true"] = np.concatenate([p["ncomp_true"][:, ::xstep] for p in parts], axis=0)
    return out


def limits_vectors(fittype, ncomp, vmin=-3.5, vmax=4.5):
    L = K.limits(fittype)
    lo = [vmin, L["sigmin"], L["Texmin"], L["taumin"]] * ncomp
    hi = [vmax, L["sigmax"], L["Texmax"], L["taumax"]] * ncomp
    return lo, hi


def _central_jacobian(fcn, x, fvec, qulim, ulim, res):
    """Drop-in for oracle.mpfit._fdjac2: central differences with a 1e-6 relative step (error ~1e-11 instead of the
    forward difference's ~1e-8) -- mpfit as it would run with an accurate Jacobian."""
    n = len(x)
    fjac = np.zeros([len(fvec), n])
    h = 1e-6 * np.abs(x)
    h[h == 0] = 1e-6
    for j in range(n):
        xp, xm = x.copy(), x.copy()
        xp[j] += h[j]
        xm[j] -= h[j]
        fjac[:, j] = (fcn(xp) - fcn(xm)) / (2 * h[j])
        res.nfev += 1
    return fjac


def oracle_fit(tile, ncomp, lo, hi, guesses, perturb=0.0, cores=None, jacobian="forward"):
    """jacobian='forward': mpfit as the reference runs it; 'central': the same algorithm with an accurate
    Jacobian (used to measure how reproducible mpfit's own answer is, tests/test_parity_floor.py)."""
    from oracle import mpfit as MP
    g = guesses if perturb == 0.0 else np.clip(guesses * (1 + perturb), np.array(lo)[:, None, None],
                                                np.array(hi)[:, None, None])
    keep = MP._fdjac2
    if jacobian == "central":
        MP._fdjac2 = _central_jacobian          # inherited by the forked workers
    try:
        return F.fiteach(tile["cube"], tile["v"], g, lo, hi, tile["fittype"],
                         multicore=cores or min(8, F.default_cores() + 1))
    finally:
        MP._fdjac2 = keep


def pixel_major(tile, guesses):
    """cube[nchan, ny, nx] -> spectra[npix, stride] (NaN padded), guesses[P, ny, nx] -> [npix, P]."""
    cube = tile["cube"]
    nchan, ny, nx = cube.shape
    npix = ny * nx
    stride = (nchan + 31) // 32 * 32
    spec = np.full((npix, stride), np.nan, dtype=np.float32)
    spec[:, :nchan] = cube.reshape(nchan, npix).T
    gg = np.ascontiguousarray(guesses.reshape(guesses.shape[0], npix).T)
    return spec, gg


def criteria(par, chi2, status, oracle, P, perr=None):
    """North-star criteria against an oracle result dict.  Returns dict of boolean arrays [npix]."""
    npix = par.shape[0]
    op = oracle["parcube"].reshape(P, npix).T
    oe = oracle["errcube"].reshape(P, npix).T
    oc = oracle["chi2"].ravel()
    ost = oracle["status"].ravel()
    conv_o = np.isin(ost, [1, 2, 3, 4])
    conv_g = np.isin(status, [1, 2, 3, 4, 6])   # 6 = converged to the precision of the fp32 evaluation (MINPACK info 6)
    with np.errstate(all="ignore"):
        dz = np.abs(par - op) / oe
        # parameter pegged in the reference (mpfit reports error 0): the GPU value must sit on the same limit, or --
        # when the GPU left it free a hair inside the box -- lie within 0.05 of the GPU's own error of that limit
        if perr is None:
            alt = np.where(np.abs(par - op) <= 1e-6, 0.0, np.inf)
        else:
            alt = np.where(np.abs(par - op) <= 1e-6, 0.0, np.where(perr > 0, np.abs(par - op) / perr, np.inf))
        dz = np.where(oe > 0, dz, alt)
        mz = dz.max(axis=1)
        rc = np.abs(chi2 - oc) / np.maximum(oc, 1e-300)
    return dict(both=conv_o & conv_g, conv_o=conv_o, conv_g=conv_g, mz=mz, rchi2=rc,
                par_ok=mz <= 0.05, chi2_ok=rc <= 1e-4)


def stable_set(o_a, o_b, P):
    """Pixels on which the ORACLE ITSELF is reproducible: refitting with guesses scaled by (1 + 1e-9) lands on
    the same solution (0.05 sigma, 1e-4 chi2, same status).  mpfit on multi-modal / degenerate spectra is chaotic
    in its starting point; no independent implementation can match it there ("Parity")."""
    npix = o_a["chi2"].size
    c = criteria(o_b["parcube"].reshape(P, npix).T, o_b["chi2"].ravel(), o_b["status"].ravel(), o_a, P)
    return c["par_ok"] & c["chi2_ok"] & (o_a["status"].ravel() == o_b["status"].ravel())


# ---- test-only host emulation of the kernel's numerical algorithm -------------------------------------------
def build_emul():
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    lib = os.path.join(ROOT, "tests", "host_emul", "libemul.so")
    deps = [src, os.path.join(ROOT, "mufasa_b200", "csrc", "b200fit_common.h"),
            os.path.join(ROOT, "mufasa_b200", "csrc", "b200fit_host.h"), os.path.join(ROOT, "include", "b200fit.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", lib, src])
    return C.CDLL(lib)


def emul_fit(lib, prob, spec, gg, mode=0):
    """mode 0: the kernels' algorithm; 1: + fp64 polish phase; 2: fp64 throughout (measured alternatives)."""
    lib.emul_set_mode(C.c_int(mode))
    npix, P = gg.shape
    out = dict(params=np.zeros((npix, P)), perr=np.zeros((npix, P)), chi2=np.zeros(npix), err_used=np.zeros(npix),
               status=np.zeros(npix, np.int32), counts=np.zeros((npix, 4), np.uint32))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_fit(C.byref(prob), vp(spec), C.c_int64(spec.shape[1]), vp(gg), None, vp(out["params"]),
                      vp(out["perr"]), vp(out["chi2"]), vp(out["err_used"]), vp(out["status"]), vp(out["counts"]))
    assert rc == 0
    return out
