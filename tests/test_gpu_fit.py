"""GPU parity tests of the fit kernel, through the C ABI, against the CPU oracle (mpfit restatement)."""
import json
import os

import numpy as np
import pytest

import _tiles as T
from mufasa_b200 import _abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    from mufasa_b200.engine import get_engine
    assert torch.cuda.is_available()
    return get_engine(0)


def _gpu_fit(engine, tile, ncomp, lo, hi, guesses):
    import torch
    spec, gg = T.pixel_major(tile, guesses)
    prob = _abi.make_problem(tile["fittype"], ncomp, tile["nchan"], spec.shape[0], tile["v"][0], tile["dv"], lo, hi)
    out = engine.fit(prob, torch.from_numpy(spec).cuda(), torch.from_numpy(gg).cuda())
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}


CASES = [
    # cfg, ncomp, nrows, xstep
    (2, 1, 4, 8),      # NH3 800 ch
    (2, 2, 4, 4),      # 256 pixels: the pass fraction of the chaotic 2-component fits has a counting noise of ~0.025
    (4, 2, 3, 8),      # N2H+ 1200 ch, low S/N, guesses on limits
    (4, 1, 3, 16),
    (5, 2, 3, 16),     # NH3 1000 ch: the channel count of BASELINE configs[2] and configs[4]
    (5, 1, 3, 32),
]


@pytest.mark.parametrize("cfg,ncomp,nrows,xstep", CASES)
def test_fit_parity(engine, cfg, ncomp, nrows, xstep):
    """North-star criteria 1 and 2 (parameters within 0.05 sigma of the reference's errors, relative chi2 within 1e-4
    on the pixels both sides flag converged) against the CPU oracle, on ALL both-converged pixels -- next to the
    same numbers for mpfit against ITSELF with an accurate Jacobian (the floor a perfect independent implementation
    can expect: 2-component fits of spectra holding fewer components are chaotic in mpfit, tests/test_parity_floor.py)
    and on the pixels where mpfit is reproducible.  Every miss goes to gpurun_out/parity_misses.jsonl."""
    tile = T.sample_tile(cfg, nrows=nrows, xstep=xstep)
    lo, hi = T.limits_vectors(tile["fittype"], ncomp)
    g = synth.clamp_guesses(tile["guess%d" % ncomp], lo, hi)
    P = 4 * ncomp
    o = T.oracle_fit(tile, ncomp, lo, hi, g)
    npix = o["chi2"].size
    o2 = T.oracle_fit(tile, ncomp, lo, hi, g, perturb=1e-9)
    stable = T.stable_set(o, o2, P)
    oc = T.oracle_fit(tile, ncomp, lo, hi, g, jacobian="central")
    c_floor = T.criteria(oc["parcube"].reshape(P, npix).T, oc["chi2"].ravel(), oc["status"].ravel(), o, P,
                         perr=oc["errcube"].reshape(P, npix).T)
    floor_ok = (c_floor["par_ok"] & c_floor["chi2_ok"])[c_floor["both"]]
    out = _gpu_fit(engine, tile, ncomp, lo, hi, g)
    c = T.criteria(out["params"], out["chi2"], out["status"], o, P, perr=out["perr"])
    ok = c["par_ok"] & c["chi2_ok"]
    both, sel = c["both"], c["both"] & stable
    miss = both & ~ok
    sg = (out["chi2"] - o["chi2"].ravel()) / np.maximum(o["chi2"].ravel(), 1e-300)
    rep = dict(cfg=cfg, ncomp=ncomp, nchan=int(tile["nchan"]), npix=int(npix),
               oracle_converged=int(c["conv_o"].sum()), gpu_converged=int(c["conv_g"].sum()),
               both_converged=int(both.sum()), passed=int(ok[both].sum()),
               oracle_self_both=int(c_floor["both"].sum()), oracle_self_passed=int(floor_ok.sum()),
               oracle_stable=int((stable & c["conv_o"]).sum()), compared_stable=int(sel.sum()),
               passed_stable=int(ok[sel].sum()), median_dsigma=float(np.median(c["mz"][both])),
               misses=int(miss.sum()), miss_gpu_chi2_lower=int((sg[miss] < -1e-4).sum()),
               miss_chi2_equal=int((np.abs(sg[miss]) <= 1e-4).sum()), miss_oracle_chi2_lower=int((sg[miss] > 1e-4).sum()),
               miss_status6=int((out["status"][miss] == 6).sum()), status6_all=int((out["status"] == 6).sum()),
               miss_truth_ncomp_below_fit=int((tile["ncomp_true"].ravel()[miss] < ncomp).sum()),
               evals_mean=float(out["counts"][:, 0].mean()),
               err_used_rel=float(np.nanmax(np.abs(out["err_used"] - o["err"].ravel()) / o["err"].ravel())))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_fit.jsonl", "a") as f:
        f.write(json.dumps(rep) + "\n")
    op = o["parcube"].reshape(P, npix).T
    oe = o["errcube"].reshape(P, npix).T
    with open("gpurun_out/parity_misses.jsonl", "a") as f:
        for i in np.nonzero(miss)[0]:
            f.write(json.dumps(dict(cfg=cfg, ncomp=ncomp, pixel=int(i), truth_ncomp=int(tile["ncomp_true"].ravel()[i]),
                                    status_oracle=int(o["status"].ravel()[i]), status_gpu=int(out["status"][i]),
                                    oracle_stable=bool(stable[i]), dchi2_rel=float(sg[i]),
                                    lower_chi2="gpu" if sg[i] < -1e-4 else ("oracle" if sg[i] > 1e-4 else "equal"),
                                    max_dsigma=float(c["mz"][i]),
                                    pegged_oracle=[int(k) for k in np.nonzero(oe[i] == 0)[0]],
                                    pegged_gpu=[int(k) for k in np.nonzero(out["perr"][i] == 0)[0]],
                                    p_oracle=[float(x) for x in op[i]], p_gpu=[float(x) for x in out["params"][i]],
                                    evals_gpu=int(out["counts"][i, 0]))) + "\n")
    print(rep)
    # has_fit agreement (status > 0 on both sides)
    assert np.array_equal(out["status"] > 0, o["status"].ravel() > 0)
    # weights: std of the spectrum (fp64 on the GPU, cube dtype in numpy) agree to fp32 round-off
    assert rep["err_used_rel"] < 5e-7
    assert both.sum() >= 0.85 * npix
    frac, floor = ok[both].mean(), floor_ok.mean()
    if ncomp == 1:
        # well-posed fits: the criteria hold on (nearly) every pixel -- as they do for mpfit against itself
        assert frac >= 0.95 and frac >= floor - 0.05
    else:
        # 2-component fits: near mpfit's own reproducibility.  The floor is mpfit with a Jacobian perturbed at the
        # 1e-8 level; the kernels differ from mpfit by an fp32 model (1e-7), an analytic Jacobian and a different
        # rounding order, so a few more of the chaotic pixels flip (counting noise of ~200 pixels: sigma ~ 0.025)
        assert frac >= floor - 0.12, "far below mpfit's own reproducibility: %.3f vs %.3f" % (frac, floor)
        # ... on the pixels where mpfit is reproducible under a 1e-9 perturbation the criteria mostly hold
        assert ok[sel].mean() >= 0.88
        # ... and the misses are not one-sided: the GPU's chi2 is the lower one about as often as mpfit's
        assert rep["miss_oracle_chi2_lower"] <= max(4, 0.75 * rep["misses"])
    # typical agreement is orders of magnitude inside the 0.05-sigma criterion
    assert rep["median_dsigma"] < 5e-3


def test_fit_matches_host_emulation(engine):
    """The kernel and the test-only host emulation run the same shared source: results agree to fp32 noise."""
    tile = T.sample_tile(2, nrows=2, xstep=8)
    lo, hi = T.limits_vectors(tile["fittype"], 2)
    g = synth.clamp_guesses(tile["guess2"], lo, hi)
    out = _gpu_fit(engine, tile, 2, lo, hi, g)
    spec, gg = T.pixel_major(tile, g)
    prob = _abi.make_problem(tile["fittype"], 2, tile["nchan"], spec.shape[0], tile["v"][0], tile["dv"], lo, hi)
    em = T.emul_fit(T.build_emul(), prob, spec, gg)
    # same outcome class: converged (1-4, or 6 = to the precision of the fp32 evaluation) / maxiter / failed.
    # Which of the ftol-type tests fires first (1 vs 6) depends on last-bit differences (ex2.approx vs exp2f).
    cls = lambda st: np.where(np.isin(st, [1, 2, 3, 4, 6]), 1, st)
    same = cls(out["status"]) == cls(em["status"])
    assert same.mean() > 0.9
    with np.errstate(all="ignore"):
        d = np.abs(out["params"] - em["params"]) / np.where(em["perr"] > 0, em["perr"], np.inf)
    assert np.median(np.nanmax(d, axis=1)[same]) < 1e-2


def test_edge_cases(engine):
    """NaN guesses, guesses outside the limits, all-NaN spectra, NaN channels, signal_cut."""
    import torch
    tile = T.sample_tile(2, nrows=1, xstep=32)
    lo, hi = T.limits_vectors(tile["fittype"], 1)
    g = synth.clamp_guesses(tile["guess1"], lo, hi)
    spec, gg = T.pixel_major(tile, g)
    npix = spec.shape[0]
    gg[0, 2] = np.nan            # NaN guess -> not fitted
    gg[1, 1] = 5.0               # sigma above its upper limit -> mpfit status 0 -> not fitted
    spec[2, :] = np.nan          # empty spectrum
    spec[3, 100:140] = np.nan    # NaN channels carry no weight, pixel still fitted
    prob = _abi.make_problem(tile["fittype"], 1, tile["nchan"], npix, tile["v"][0], tile["dv"], lo, hi)
    out = engine.fit(prob, torch.from_numpy(spec).cuda(), torch.from_numpy(gg).cuda())
    st = out["status"].cpu().numpy()
    par = out["params"].cpu().numpy()
    assert st[0] == 0 and st[1] == 0 and st[2] == 0
    assert np.all(par[:3] == 0) and np.all(out["perr"].cpu().numpy()[:3] == 0)
    assert st[3] > 0 and np.all(np.isfinite(par[3]))
    assert np.all(st[4:] > 0)
    # signal_cut: a cut above every peak S/N leaves nothing fitted (pyspeckit fiteach semantics)
    prob2 = _abi.make_problem(tile["fittype"], 1, tile["nchan"], npix, tile["v"][0], tile["dv"], lo, hi,
                              signal_cut=1e6)
    out2 = engine.fit(prob2, torch.from_numpy(spec).cuda(), torch.from_numpy(gg).cuda())
    assert np.all(out2["status"].cpu().numpy() == 0)
    # empty call
    prob0 = _abi.make_problem(tile["fittype"], 1, tile["nchan"], 0, tile["v"][0], tile["dv"], lo, hi)
    out0 = engine.fit(prob0, torch.empty((0, spec.shape[1]), dtype=torch.float32, device="cuda"),
                      torch.empty((0, 4), dtype=torch.float64, device="cuda"))
    assert out0["params"].shape[0] == 0


def test_odd_sizes_weights_and_iteration_cap(engine):
    """Pixel counts that do not fill a warp of solver groups, caller-supplied weights (weight_mode 1) and a small
    max_iter: per-pixel results must not depend on the batch they are fitted in."""
    import torch
    tile = T.sample_tile(2, nrows=2, xstep=8)
    for ncomp in (1, 2):
        lo, hi = T.limits_vectors(tile["fittype"], ncomp)
        g = synth.clamp_guesses(tile["guess%d" % ncomp], lo, hi)
        spec, gg = T.pixel_major(tile, g)
        d_spec, d_g = torch.from_numpy(spec).cuda(), torch.from_numpy(gg).cuda()
        full = engine.fit(_abi.make_problem(tile["fittype"], ncomp, tile["nchan"], spec.shape[0], tile["v"][0],
                                            tile["dv"], lo, hi), d_spec, d_g)
        full = {k: v.clone() for k, v in full.items()}
        for n in (1, 3, 5, 37):
            sub = engine.fit(_abi.make_problem(tile["fittype"], ncomp, tile["nchan"], n, tile["v"][0], tile["dv"], lo,
                                               hi), d_spec[:n].contiguous(), d_g[:n].contiguous())
            for k in ("params", "perr", "chi2", "status"):
                assert torch.equal(sub[k], full[k][:n]), (ncomp, n, k)
        # weight_mode 1: err given by the caller; 2x the weights used before -> same parameters, errors x2, chi2 / 4
        prob_w = _abi.make_problem(tile["fittype"], ncomp, tile["nchan"], spec.shape[0], tile["v"][0], tile["dv"], lo,
                                   hi, weight_mode=1)
        w = engine.fit(prob_w, d_spec, d_g, err=(2.0 * full["err_used"]).contiguous())
        ok = full["status"] > 0
        assert torch.equal(w["status"], full["status"])
        assert torch.equal(w["params"][ok], full["params"][ok])
        assert torch.allclose(w["perr"][ok], 2.0 * full["perr"][ok], rtol=1e-12, atol=0)
        assert torch.allclose(w["chi2"][ok], full["chi2"][ok] / 4.0, rtol=1e-12, atol=0)
        # iteration cap: at most max_iter accepted steps, flagged 5 where it bites, still a valid fit (has_fit)
        prob_c = _abi.make_problem(tile["fittype"], ncomp, tile["nchan"], spec.shape[0], tile["v"][0], tile["dv"], lo,
                                   hi, max_iter=3)
        c = engine.fit(prob_c, d_spec, d_g)
        st = c["status"].cpu().numpy()
        assert np.all(st > 0) and (st == 5).any()
        assert int(c["counts"][:, 0].max()) <= 2 * 3 + 2


def test_descending_velocity_axis():
    """A cube stored with descending velocity (common in FITS) gives the same fit as its ascending twin."""
    from mufasa_b200.cube import SpecCube
    from mufasa_b200.UltraCube import UltraCube
    syn = synth.make_cube(2, T.oracle_modelcube, rows=(100, 102), shape=(200, 24, 800))
    res = []
    for flip in (False, True):
        cube, v = (syn["cube"][::-1].copy(), syn["v"][::-1].copy()) if flip else (syn["cube"], syn["v"])
        uc = UltraCube(cube=SpecCube(cube, v), fittype=syn["fittype"], device=0)
        uc.fit_cube(ncomp=1, simpfit=True, guesses=syn["guess1"].copy())
        res.append(uc.pcubes["1"])
    assert np.array_equal(res[0].has_fit, res[1].has_fit)
    assert np.array_equal(res[0].parcube, res[1].parcube)
    assert np.array_equal(res[0].errcube, res[1].errcube)
    m0, m1 = res[0].get_modelcube(), res[1].get_modelcube()
    assert np.array_equal(m0, m1[::-1], equal_nan=True)
