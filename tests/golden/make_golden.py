#!/usr/bin/env python
"""Generate the committed golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box never runs this; tests read the committed ``*.npz`` / ``*.json`` next to this file.

What is executed from /root/reference (through the import stubs of _ref_stubs.py):
  model_golden.npz   mufasa/spec_models/HyperfineModel.py:22-109 (AmmoniaModel / N2HplusModel .multi_v_spectrum)
  aic_golden.npz     mufasa/aic.py:136-201 (AIC, AICc, likelihood)
  stats_golden.npz   mufasa/UltraCube.py: expand_mask :1658, get_rms :1701, get_rss :1384, get_chisq :1478,
                     and -- through a real ``UltraCube`` instance holding stand-in pcubes --
                     get_all_lnk_maps :1218, get_best_2c_parcube :1283 (=> get_AICc :474, get_rss :432,
                     calc_rss :911, calc_AICc_likelihood :1120, reset_model_mask :359, has_fit :328)
  guess_golden.npz   mufasa/moment_guess.py: moment_guesses :388, moment_guesses_1c :523, get_tex/get_tau/peakT
                     :774-800;
                     mufasa/guess_refine.py: tautex_renorm :183
  meta_golden.json   mufasa/spec_models/meta_model.py:67-150 (limits, windows, rest value, voff_at_peak)
Not reference-provided: the NH3 (1,1) hyperfine table (pyspeckit, absent) -- injected, see _ref_stubs.py.
The LM solver (pyspeckit/mpfit) cannot be executed here; no golden exists for it ("parity unpinned").
"""
import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import _ref_stubs as S  # noqa: E402

S.install()

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

from mufasa import UltraCube as UC  # noqa: E402
from mufasa import aic  # noqa: E402
from mufasa import guess_refine as gr  # noqa: E402
from mufasa import moment_guess as mg  # noqa: E402
from mufasa.spec_models.SpecModels import AmmoniaModel, N2HplusModel  # noqa: E402
from mufasa.spec_models.meta_model import MetaModel  # noqa: E402

CKMS = 299792.458


def vaxis(nchan, dv):
    return (np.arange(nchan) - nchan // 2) * dv


def ghz(v, nu0_hz):
    return nu0_hz * (1.0 - v / CKMS) / 1e9


def make_model():
    out = {}
    rng = np.random.default_rng(20261019)
    for key, cls, nu0, nchan, dv, taumax in (("nh3", AmmoniaModel, 23.6944955e9, 500, 0.10, 30.0),
                                             ("n2hp", N2HplusModel, 93173.7637e6, 600, 0.10, 40.0)):
        mod = cls()
        v = vaxis(nchan, dv)
        x = S.FakeXarr(ghz(v, nu0))
        pars1 = [[0.0, 0.2, 6.0, 2.0], [-1.7, 0.07, 3.0, 0.1], [2.3, 2.5, 40.0, taumax], [0.61, 0.33, 4.1, 11.0]]
        pars2 = [[0.3, 0.25, 6.5, 3.0, 1.4, 0.2, 5.2, 1.2], [-0.5, 0.12, 3.4, 0.4, -0.45, 0.9, 9.0, 7.0]]
        for _ in range(4):
            p = [rng.uniform(-2, 2), rng.uniform(0.07, 1.0), rng.uniform(3, 12), rng.uniform(0.1, 10)]
            pars1.append(p)
            q = [p[0] + rng.uniform(0.2, 1.6), rng.uniform(0.07, 0.6), rng.uniform(3, 9), rng.uniform(0.1, 5)]
            pars2.append(p + q)
        out[key + "_v"] = v
        out[key + "_nu0"] = np.float64(nu0)
        out[key + "_pars1"] = np.array(pars1)
        out[key + "_pars2"] = np.array(pars2)
        out[key + "_spec1"] = np.array([mod.multi_v_spectrum(x, *p) for p in pars1])
        out[key + "_spec2"] = np.array([mod.multi_v_spectrum(x, *p) for p in pars2])
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)


def make_aic():
    rng = np.random.default_rng(7)
    rss = rng.uniform(0.5, 50.0, size=(6, 7))
    N = rng.integers(0, 400, size=(6, 7)).astype(float)
    N[0, 0] = 0.0
    N[1, 1] = 9.0          # N - p - 1 == 0 for p = 8
    N[2, 2] = 5.0
    out = dict(rss=rss, N=N)
    for p in (0, 4, 8):
        out["AIC_p%d" % p] = aic.AIC(rss.copy(), p, N.copy())
        out["AICc_p%d" % p] = aic.AICc(rss.copy(), p, N.copy())
    out["lnk"] = aic.likelihood(out["AICc_p8"], out["AICc_p4"])
    np.savez_compressed(os.path.join(HERE, "aic_golden.npz"), **out)


class _FakePCube(object):
    def __init__(self, parcube, errcube, modelcube):
        self.parcube = parcube
        self.errcube = errcube
        self._modelcube = modelcube

    def get_modelcube(self, update=False, multicore=1):
        return self._modelcube


class _FakeSCube(S.FakeCube):
    spectral_axis = None


def make_stats():
    rng = np.random.default_rng(11)
    nchan, ny, nx, dv = 240, 5, 6, 0.2
    nu0 = 23.6944955e9
    v = vaxis(nchan, dv)
    x = S.FakeXarr(ghz(v, nu0))
    mod = AmmoniaModel()
    par1 = np.zeros((4, ny, nx))
    par2 = np.zeros((8, ny, nx))
    err1 = np.zeros((4, ny, nx))
    err2 = np.zeros((8, ny, nx))
    data = np.zeros((nchan, ny, nx))
    for y in range(ny):
        for xx in range(nx):
            p = [rng.uniform(-1, 1), rng.uniform(0.15, 0.5), rng.uniform(4, 9), rng.uniform(0.5, 6)]
            q = [p[0] + rng.uniform(0.5, 1.6), 0.8 * p[1], 0.8 * p[2], 0.4 * p[3]]
            two = (xx % 2 == 0)
            truth = p + q if two else p
            data[:, y, xx] = mod.multi_v_spectrum(x, *truth) + rng.normal(0, 0.1, nchan)
            par1[:, y, xx] = np.array(p) * (1 + 0.01 * rng.normal(size=4))
            par2[:, y, xx] = np.array(p + q) * (1 + 0.01 * rng.normal(size=8))
            err1[:, y, xx] = rng.uniform(0.01, 0.1, 4)
            err2[:, y, xx] = rng.uniform(0.01, 0.1, 8)
    # pixels without fits (all zero = "no fit", UltraCube.py:345-348), one noise-only pixel, NaN channels
    par1[:, 0, 0] = 0
    err1[:, 0, 0] = 0
    par2[:, 0, 0] = 0
    err2[:, 0, 0] = 0
    par2[:, 1, 1] = 0
    err2[:, 1, 1] = 0
    data[:, 4, 5] = rng.normal(0, 0.1, nchan)
    data[10:14, 2, 2] = np.nan
    data[120:123, 3, 3] = np.nan
    data = data.astype(np.float32)

    def modelcube(par, dtype):
        out = np.full((nchan, ny, nx), np.nan, dtype=dtype)
        for y in range(ny):
            for xx in range(nx):
                if np.any(par[:, y, xx] != 0):
                    out[:, y, xx] = mod.multi_v_spectrum(x, *par[:, y, xx])
        return out

    out = dict(v=v, nu0=np.float64(nu0), data=data, par1=par1, par2=par2, err1=err1, err2=err2)
    for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
        m1 = modelcube(par1, dtype)
        m2 = modelcube(par2, dtype)
        out["model1_" + tag] = m1.copy()
        out["model2_" + tag] = m2.copy()
        # module-level statistics on the raw model cube of the 2-comp fit
        mask = m2 > 0
        out["expand20_" + tag] = UC.expand_mask(mask.copy(), 20)
        out["expand7_" + tag] = UC.expand_mask(mask.copy(), 7)
        cube = _FakeSCube(data)
        rss, ns = UC.get_rss(cube, m2.copy(), expand=20, usemask=True, mask=None)
        out["rss_nomask_" + tag], out["ns_nomask_" + tag] = rss, ns
        rss, ns = UC.get_rss(cube, m1.copy(), expand=0, usemask=True, mask=(m1 > 0) | (m2 > 0))
        out["rss_e0_" + tag], out["ns_e0_" + tag] = rss, ns
        resid = data - np.nan_to_num(m2)
        out["rms_" + tag] = UC.get_rms(resid)
        out["chisq_red_" + tag] = UC.get_chisq(cube, np.nan_to_num(m2), expand=20, reduced=True)
        # class-level model selection through the reference's UltraCube object
        uc = UC.UltraCube(cube=cube, fittype=None, n_cores=1)
        uc.pcubes["1"] = _FakePCube(par1.copy(), err1.copy(), m1.copy())
        uc.pcubes["2"] = _FakePCube(par2.copy(), err2.copy(), m2.copy())
        parcube, errcube, lnk10, lnk20, lnk21 = uc.get_best_2c_parcube(multicore=1)
        out["best_par_" + tag], out["best_err_" + tag] = parcube, errcube
        out["lnk10_" + tag], out["lnk20_" + tag], out["lnk21_" + tag] = lnk10, lnk20, lnk21
        for nc in ("0", "1", "2"):
            out["rssmap%s_%s" % (nc, tag)] = uc.rss_maps[nc]
            out["nsamp%s_%s" % (nc, tag)] = uc.NSamp_maps[nc]
            out["aicc%s_%s" % (nc, tag)] = uc.AICc_maps[nc]
        out["master_mask_" + tag] = uc.master_model_mask
    np.savez_compressed(os.path.join(HERE, "stats_golden.npz"), **out)


def make_guess():
    rng = np.random.default_rng(3)
    m1 = rng.uniform(-2, 2, (7, 9))
    m2 = rng.uniform(0.01, 1.2, (7, 9))
    m0 = rng.uniform(0.1, 12.0, (7, 9))
    m0[0, 0] = np.nan
    m1[0, 0] = np.nan
    out = dict(m0=m0, m1=m1, m2=m2)
    for nc in (1, 2):
        out["gg%d_nh3" % nc] = mg.moment_guesses(m1.copy(), m2.copy(), nc, sigmin=0.07, moment0=m0.copy(),
                                                linetype='nh3', mom0_floor=None)
        # (moment0=None makes the reference raise TypeError at moment_guess.py:513 -- scalar tex_guess is
        #  item-assigned; cubefit_gen always passes moment0, multi_v_fit.py:297)
        out["gg%d_n2hp" % nc] = mg.moment_guesses(m1.copy(), m2.copy(), nc, sigmin=0.05, moment0=m0.copy(),
                                                 linetype='n2hp', mom0_floor=0.5)
    tau = rng.uniform(0.01, 6.0, (2, 7, 9))
    tex = rng.uniform(3.0, 30.0, (2, 7, 9))
    tau[0, 1, 1] = np.nan
    out["renorm_tau_in"], out["renorm_tex_in"] = tau.copy(), tex.copy()
    for key, nu in (("nh3", 23.6944955), ("n2hp", 93.1737637)):
        t, x = gr.tautex_renorm(tau.copy(), tex.copy(), tau_thresh=0.1, nu=nu)
        out["renorm_tau_%s" % key], out["renorm_tex_%s" % key] = t, x
    t, x = gr.tautex_renorm(tau.copy(), tex.copy())          # defaults (tau_thresh=0.21, nu0_nh3)
    out["renorm_tau_default"], out["renorm_tex_default"] = t, x
    Ta = rng.uniform(0.05, 4.0, 20)
    out["Ta"] = Ta
    out["get_tex"] = mg.get_tex(Ta, tau=0.7, nu=23.6944955)
    out["get_tau"] = mg.get_tau(Ta * 0.3, tex=6.0, nu=23.6944955)
    out["peakT"] = mg.peakT(rng.uniform(3, 12, 20) * 0 + np.linspace(3, 12, 20), np.linspace(0.1, 5, 20), nu=93.1737637)
    # moment_guesses_1c (moment_guess.py:523-572): amplitude-like moment 0 spanning both tau/Tex regimes, the floor
    # and a few NaNs
    rng1 = np.random.default_rng(77)
    m0a = rng1.uniform(0.0, 6.0, size=(5, 7))
    m0a[0, 0], m0a[4, 6] = 1e-4, np.nan
    m1a = rng1.uniform(-3, 3, size=(5, 7))
    m2a = rng1.uniform(0.05, 1.5, size=(5, 7))
    out["m1c_m0"], out["m1c_m1"], out["m1c_m2"] = m0a, m1a, m2a
    out["m1c_out"] = mg.moment_guesses_1c(m0a.copy(), m1a.copy(), m2a.copy())
    np.savez_compressed(os.path.join(HERE, "guess_golden.npz"), **out)


def make_cnv():
    """Pure-numpy pieces of the convolved-cube stage: component sorting / swapping (guess_refine.py:36-110),
    refine_2c_guess (:395-416), downsample_header (utils/fits_utils.py:11-60)."""
    from mufasa.utils import fits_utils as fu
    rng = np.random.default_rng(11)
    data = rng.uniform(0.1, 9.0, (16, 6, 7))
    data[:, 0, 0] = np.nan
    data[8, 1, 1] = np.nan
    out = dict(data=data)
    swapmask = rng.uniform(size=(6, 7)) > 0.5
    out["swapmask"] = swapmask
    out["swapped"] = gr.mask_swap_2comp(data.copy(), swapmask)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in ("error_v", "Tpeak", "tautex", "tau", "chen2020"):
            out["sort_" + m] = gr.quick_2comp_sort(data.copy(), filtsize=2, method=m)
        out["refine_2c"] = gr.refine_2c_guess(data[:8].copy(), f_sigv=0.5)
    rows = []
    for crpix, cdelt, factor in ((50.0, 0.1, 2), (128.5, -0.002444, 2), (1.0, 0.01, 3), (33.25, -1.5e-3, 2.5)):
        h = {"CRPIX1": crpix, "CDELT1": cdelt, "CRPIX2": crpix + 3, "CDELT2": -cdelt}
        h1 = fu.downsample_header(h, factor=factor, axis=1)
        h2 = fu.downsample_header(h1, factor=factor, axis=2)
        rows.append([crpix, cdelt, factor, h2["CRPIX1"], h2["CDELT1"], h2["CRPIX2"], h2["CDELT2"]])
    out["downsample_header"] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "cnv_golden.npz"), **out)


def make_meta():
    out = {}
    for ft in ("nh3_multi_v", "n2hp_multi_v"):
        m = MetaModel(ft, 2)
        out[ft] = dict(Texmin=m.Texmin, Texmax=m.Texmax, sigmin=m.sigmin, sigmax=m.sigmax, taumin=m.taumin,
                       taumax=m.taumax, eps=m.eps, window_hwidth=m.window_hwidth,
                       central_win_hwidth=m.central_win_hwidth, linetype=m.linetype, linename=m.linename,
                       rest_value_hz=float(m.rest_value.value), voff_at_peak=float(m.voff_at_peak),
                       voffs=[float(a) for a in m.voffs], tau_wts=[float(a) for a in m.tau_wts])
    try:
        MetaModel("bogus", 1)
        out["bogus_raises"] = None
    except Exception as e:                                   # FitTypeError (LookupError), meta_model.py:135
        out["bogus_raises"] = [type(e).__name__, [b.__name__ for b in type(e).__mro__]]
    with open(os.path.join(HERE, "meta_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    make_model()
    make_aic()
    make_stats()
    make_guess()
    make_cnv()
    make_meta()
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith((".npz", ".json")):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))
