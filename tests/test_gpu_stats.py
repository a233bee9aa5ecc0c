"""GPU parity of the statistics half of the path (model cube, AICc selection, residual statistics, cube gather),
through the C ABI, against the CPU oracle -- which is itself pinned to the reference's own code by the golden
vectors (tests/test_oracle_golden.py).  Parameters come from the ORACLE's fits so that the only thing compared is
the statistics arithmetic."""
import numpy as np
import pytest

import _tiles as T
from mufasa_b200 import _abi, synth
from oracle import model as M
from oracle import stats as ST

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    from mufasa_b200.engine import get_engine
    assert torch.cuda.is_available()
    return get_engine(0)


@pytest.fixture(scope="module", params=[2, 4])
def fitted_tile(request):
    """A sampled tile of config 2 (NH3) / 4 (N2H+, low SNR) with the oracle's 1c and 2c fits; a few pixels are
    blanked (no fit) and one keeps a fit whose model never rises above zero (empty mask -> include_nosamp)."""
    cfg = request.param
    tile = T.sample_tile(cfg, nrows=2, xstep=16 if cfg == 2 else 32)
    fits = {}
    for nc in (1, 2):
        lo, hi = T.limits_vectors(tile["fittype"], nc)
        g = synth.clamp_guesses(tile["guess%d" % nc], lo, hi)
        fits[nc] = T.oracle_fit(tile, nc, lo, hi, g)
    p1, p2 = fits[1]["parcube"].copy(), fits[2]["parcube"].copy()
    p1[:, 0, 1] = 0.0                       # no 1c fit
    p2[:, 0, 2] = 0.0                       # no 2c fit
    p1[:, 1, 0] = 0.0
    p2[:, 1, 0] = 0.0                       # neither
    tile["p1"], tile["p2"] = p1, p2
    return tile


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _prob(tile, ncomp, npix):
    P = 4 * ncomp
    return _abi.make_problem(tile["fittype"], ncomp, tile["nchan"], npix, tile["v"][0], tile["dv"], [-1e30] * P,
                             [1e30] * P)


def test_gather_spectra(engine):
    import torch
    rng = np.random.default_rng(5)
    nchan, ny, nx = 70, 5, 9
    cube = rng.normal(size=(nchan, ny, nx)).astype(np.float32)
    cube[3, 2, 4] = np.nan
    d = _dev(cube.reshape(nchan, ny * nx))
    rows = engine.gather(d).cpu().numpy()
    assert rows.shape == (ny * nx, 96)
    assert np.array_equal(rows[:, :nchan], cube.reshape(nchan, -1).T, equal_nan=True)
    assert np.all(np.isnan(rows[:, nchan:]))
    pix = np.array([44, 3, 17], dtype=np.int64)
    rows = engine.gather(d, torch.from_numpy(pix).cuda(), flip=True).cpu().numpy()
    assert np.array_equal(rows[:, :nchan], cube.reshape(nchan, -1).T[pix][:, ::-1], equal_nan=True)


def test_modelcube(engine, fitted_tile):
    t = fitted_tile
    for nc, par in ((1, t["p1"]), (2, t["p2"])):
        P, ny, nx = par.shape
        ref = M.modelcube(t["v"], par, t["fittype"])                      # fp64, reference operation order
        out = engine.model(_prob(t, nc, ny * nx), _dev(par.reshape(P, -1).T)).cpu().numpy()[:, :t["nchan"]]
        out = out.T.reshape(t["nchan"], ny, nx)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        ok = ~np.isnan(ref)
        # same operations in the same order; CUDA's exp() is within 1 ulp of glibc's -> 1e-14 of the model scale
        assert np.max(np.abs(out[ok] - ref[ok])) <= 1e-13 * np.max(np.abs(ref[ok]))
        # the AICc sample mask is ``model > 0``: decided by fp64 round-off at the window edges
        agree = ((out > 0) == (ref > 0))[ok].mean()
        assert agree > 0.998, agree    # ~1 edge channel per spectrum sits at 1 - exp(-tau) ~ 1e-16


def test_aicc_selection(engine, fitted_tile):
    t = fitted_tile
    cube = t["cube"]
    nchan, ny, nx = cube.shape
    npix = ny * nx
    m1 = M.modelcube(t["v"], t["p1"], t["fittype"], dtype=np.float32)
    m2 = M.modelcube(t["v"], t["p2"], t["fittype"], dtype=np.float32)
    with np.errstate(all="ignore"):
        L = ST.all_lnk_maps(cube, m1, m2, expand=20)
        e1, e2 = np.zeros_like(t["p1"]), np.zeros_like(t["p2"])
        _, _, l10, l20, l21, ncomp = ST.best_2c_parcube(t["p1"], e1, t["p2"], e2, L["lnk10"], L["lnk20"], L["lnk21"])
    spec, _ = T.pixel_major(t, t["p1"])
    out = engine.aicc_global(_prob(t, 1, npix), _dev(spec), _dev(t["p1"].reshape(4, -1).T),
                             _dev(t["p2"].reshape(8, -1).T), expand=20, model_f32=True)
    g = {k: v.cpu().numpy() for k, v in out.items() if hasattr(v, "cpu")}
    nsamp = g["nsamp"].reshape(ny, nx)
    # N: integer counts.  The channel at each edge of a hyperfine window has 1 - exp(-tau) ~ 1e-16: whether it
    # belongs to ``model > 0`` depends on the last bit of exp() -- CUDA's, glibc's and numpy's SIMD exp differ
    # there, i.e. the reference's own N is platform dependent by +-1..2 (measured: 30 % of the pixels, always by
    # <= 2 channels of ~400).  Everything downstream is compared exactly where N agrees and loosely elsewhere.
    both = np.isfinite(nsamp) & np.isfinite(L["nsamp"])
    assert np.array_equal(np.isfinite(nsamp), np.isfinite(L["nsamp"]))
    dn = np.abs(nsamp[both] - L["nsamp"][both])
    assert np.mean(dn == 0) > 0.5 and dn.max() <= 2
    exact = both & (nsamp == L["nsamp"])
    for k, name in enumerate(("rss0", "rss1", "rss2")):
        r = g["rss"].reshape(ny, nx, 3)[:, :, k]
        assert np.array_equal(np.isnan(r), np.isnan(L[name]))
        sel = exact & np.isfinite(L[name])
        assert np.max(np.abs(r[sel] - L[name][sel]) / L[name][sel]) < 1e-6       # fp32 residuals, fp64 sums
    lnk = g["lnk"].reshape(ny, nx, 3)
    for k, ref in enumerate((l10, l20, l21)):
        assert np.array_equal(np.isnan(lnk[:, :, k]), np.isnan(ref))
        sel = exact & np.isfinite(ref)
        assert np.max(np.abs(lnk[:, :, k][sel] - ref[sel])) < 1e-4 * np.maximum(1.0, np.abs(ref[sel])).max()
    # where N differs by an edge channel the likelihoods move by ~1e-3 of |ln(rss_a / rss_b)|
    loose = both & ~exact
    for k, ref in enumerate((l10, l20, l21)):
        sel = loose & np.isfinite(ref)
        if sel.any():
            assert np.max(np.abs(lnk[:, :, k][sel] - ref[sel])) < 0.02 * np.maximum(1.0, np.abs(ref[sel])).max()
    # integer ncomp map: bit-exact where N agrees unless a likelihood sits within 1e-6 of its threshold (the north-star
    # rule); where N differs by an edge channel, unless it sits within 0.05 of the threshold
    nc_g = g["ncomp"].reshape(ny, nx)
    near6 = np.zeros((ny, nx), bool)
    near2 = np.zeros((ny, nx), bool)
    with np.errstate(invalid="ignore"):
        for ref in (L["lnk10"], L["lnk20"], L["lnk21"]):
            near6 |= np.abs(ref - 5.0) < 1e-6
            near2 |= np.abs(ref - 5.0) < 0.05
    mism = (nc_g != ncomp) & ((exact & ~near6) | (~exact & ~near2))
    assert not mism.any(), (np.argwhere(mism), nc_g[mism], ncomp[mism])
    assert (nc_g == ncomp).mean() > 0.97


def test_resid_stats(engine, fitted_tile):
    t = fitted_tile
    cube = t["cube"]
    nchan, ny, nx = cube.shape
    spec, _ = T.pixel_major(t, t["p1"])
    for nc, par in ((1, t["p1"]), (2, t["p2"])):
        m = M.modelcube(t["v"], par, t["fittype"], dtype=np.float32)
        with np.errstate(all="ignore"):
            rms_ref = ST.get_rms(cube - m)
            chi_ref = ST.get_chisq(cube, m.copy(), expand=20, reduced=True)
            chi_ref = chi_ref[0] if isinstance(chi_ref, tuple) else chi_ref
        out = engine.resid_stats(_prob(t, nc, ny * nx), _dev(spec), _dev(par.reshape(4 * nc, -1).T), expand=20,
                                 model_f32=True)
        rms = out["rms"].cpu().numpy().reshape(ny, nx)
        chi = out["chisq_red"].cpu().numpy().reshape(ny, nx)
        assert np.array_equal(np.isnan(rms), np.isnan(rms_ref))
        ok = np.isfinite(rms_ref)
        assert np.max(np.abs(rms[ok] - rms_ref[ok]) / rms_ref[ok]) < 2e-6       # float32 medians of |diff|
        okc = np.isfinite(chi_ref) & np.isfinite(chi)
        assert okc.sum() >= 0.8 * ok.sum()
        rel = np.abs(chi[okc] - chi_ref[okc]) / chi_ref[okc]
        # bit-level agreement where the dilated masks agree; a mask-edge channel (see test_aicc_selection) moves
        # N by 1-2 of ~400 elsewhere
        assert np.median(rel) < 1e-9 and rel.max() < 2e-2


def test_rms_snr_matches_numpy_signals(engine):
    """b200fit_rms_snr vs the numpy restatement of signals.get_rms_robust / get_snr / mad_std (oracle/signals.py):
    bit-identical (fp64 medians of the same values), including NaN channels, a fully masked pixel and a planemask."""
    import torch
    from oracle import signals
    tile = T.sample_tile(2, nrows=2, xstep=16)
    cube = tile["cube"].copy()
    cube[100:130, 0, 3] = np.nan
    cube[:, 1, 5] = np.nan
    nchan, ny, nx = cube.shape
    planemask = np.ones((ny, nx), bool)
    planemask[0, :2] = False
    include = np.isfinite(cube) & planemask[None]
    with np.errstate(all="ignore"):
        rms0_ref = signals.mad_std(cube, axis=0)                                   # snr_estimate's errmap
        rms_ref = signals.get_rms_robust(cube, sigma_cut=2.5, expand=20, include=include)
        snr_ref = signals.get_snr(cube, rms=rms_ref, include=include)
    t = dict(tile)
    t["cube"] = cube
    spec, _ = T.pixel_major(t, tile["guess1"])
    d_spec = torch.from_numpy(spec).cuda()
    a = engine.rms_snr(nchan, d_spec)
    b = engine.rms_snr(nchan, d_spec, planemask=torch.from_numpy(planemask.astype(np.uint8).ravel()).cuda())
    assert np.array_equal(a["rms0"].cpu().numpy().reshape(ny, nx), rms0_ref, equal_nan=True)
    rms = b["rms"].cpu().numpy().reshape(ny, nx)
    assert np.array_equal(rms, rms_ref, equal_nan=True)
    with np.errstate(all="ignore"):
        snr = b["peak"].cpu().numpy().reshape(ny, nx) / rms
    assert np.array_equal(snr, snr_ref, equal_nan=True)
    assert np.isnan(rms[1, 5]) and np.all(np.isnan(rms[0, :2])) and np.isfinite(rms[0, 3])


def test_aicc_pair_new_vs_old_fit(engine, fitted_tile):
    """calc_AICc_likelihood(ucube_new, nc, nc, ucube_B=old) (UltraCube.py:1120-1216, expand = 0): two 2-component
    fits of the same spectra compared over their common mask -- b200fit_aicc_pair vs the oracle's get_rss / AICc."""
    t = fitted_tile
    cube = t["cube"]
    nchan, ny, nx = cube.shape
    npix = ny * nx
    pB = t["p2"]                                  # "old" fit
    pA = pB.copy()                                # "new" fit: perturbed parameters, a few pixels without a fit
    rng = np.random.default_rng(11)
    pA[0] += 0.05 * rng.normal(size=(ny, nx))
    pA[3] *= 1.0 + 0.1 * rng.normal(size=(ny, nx))
    pA[:, 1, 2] = 0.0
    mA = M.modelcube(t["v"], pA, t["fittype"], dtype=np.float32)
    mB = M.modelcube(t["v"], pB, t["fittype"], dtype=np.float32)
    with np.errstate(all="ignore"):
        mask = np.logical_or(mA > 0, mB > 0)
        rssA, nA = ST.rss_maps(cube, mA.copy(), mask.copy(), expand=0)
        rssB, nB = ST.rss_maps(cube, mB.copy(), mask.copy(), expand=0)
        ref = ST.likelihood(ST.AICc(rssA, 8, nA), ST.AICc(rssB, 8, nB))
    spec, _ = T.pixel_major(t, t["p1"])
    out = engine.aicc_global(_prob(t, 1, npix), _dev(spec), _dev(pA.reshape(8, -1).T), _dev(pB.reshape(8, -1).T),
                             expand=0, model_f32=True, ncomp_pair=(2, 2))
    rss = out["rss"].cpu().numpy().reshape(ny, nx, 3)
    ns = out["nsamp"].cpu().numpy().reshape(ny, nx)
    with np.errstate(all="ignore"):
        nsA = np.where(np.isnan(rss[:, :, 1]), np.nan, ns)
        nsB = np.where(np.isnan(rss[:, :, 2]), np.nan, ns)
        got = ST.likelihood(ST.AICc(rss[:, :, 1], 8, nsA), ST.AICc(rss[:, :, 2], 8, nsB))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    same_n = np.isfinite(ref) & (nsA == nA)
    assert same_n.sum() > 0.5 * np.isfinite(ref).sum()
    assert np.max(np.abs(got[same_n] - ref[same_n])) < 1e-4 * np.maximum(1.0, np.abs(ref[same_n])).max()
    # the decision replace_bad_pix takes (lnk > 0) agrees wherever it is not marginal
    dec = np.isfinite(ref) & (np.abs(ref) > 0.05)
    assert np.array_equal(got[dec] > 0, ref[dec] > 0)
    # kernel's own lnk(B, A) = -lnk(A, B)
    lnkBA = out["lnk"].cpu().numpy().reshape(ny, nx, 3)[:, :, 2]
    ok = np.isfinite(lnkBA) & np.isfinite(got)
    assert np.allclose(lnkBA[ok], -got[ok], rtol=1e-10, atol=1e-10)


def test_custom_mask_rss_and_aicc_exact(fitted_tile):
    """UltraCube.get_rss(ncomp, mask=..., planemask=...) / get_AICc(..., mask=...) (UltraCube.py:432-451, :474-486)
    with a caller-supplied [nchan, ny, nx] mask: the mask is given, so N is exact and rss / AICc must agree with
    the oracle's get_rss / AICc to rounding -- including pixels whose mask is empty (include_nosamp fill from the
    pixel holding the most samples), the dilation, and the partial update inside a planemask."""
    from mufasa_b200.cube import SpecCube
    from mufasa_b200.UltraCube import UltraCube
    t = fitted_tile
    cube = t["cube"]
    nchan, ny, nx = cube.shape
    uc = UltraCube(cube=SpecCube(cube, t["v"]), fittype=t["fittype"], device=0)
    for nc, key in ((1, "p1"), (2, "p2")):
        p = uc.load_pcube()
        p.parcube, p.errcube = t[key].copy(), np.zeros_like(t[key])
        p.has_fit = np.any(p.parcube != 0, axis=0)
        p.fittype, p.npars = t["fittype"], 4 * nc
        uc.pcubes[str(nc)] = p
        uc._include_fit_in_mask(nc)
    rng = np.random.default_rng(5)
    mask = np.zeros((nchan, ny, nx), bool)
    for y in range(ny):
        for x in range(nx):
            c0 = int(rng.integers(50, nchan - 150))
            mask[c0:c0 + int(rng.integers(5, 90)), y, x] = True
    mask[:, 0, 1] = False                                    # empty masks: filled from the max-N pixel
    mask[:, 1, 3] = False
    mask[0:3, 0, 0] = True                                   # dilation is clipped at the band edge
    m1 = M.modelcube(t["v"], t["p1"], t["fittype"], dtype=np.float32)
    m2 = M.modelcube(t["v"], t["p2"], t["fittype"], dtype=np.float32)
    for expand in (20, 0):
        with np.errstate(all="ignore"):
            ref = {0: ST.rss_maps(cube, np.zeros(cube.shape), mask.copy(), expand=expand),
                   1: ST.rss_maps(cube, m1.copy(), mask.copy(), expand=expand),
                   2: ST.rss_maps(cube, m2.copy(), mask.copy(), expand=expand)}
        for nc in (0, 1, 2):
            uc._invalidate_stats()
            rss = uc.get_rss(nc, mask=mask.copy(), expand=expand)
            assert np.array_equal(uc.NSamp_maps[str(nc)], ref[nc][1], equal_nan=True), (expand, nc)
            np.testing.assert_allclose(rss, ref[nc][0], rtol=1e-6, equal_nan=True)
            aicc = uc.get_AICc(nc, mask=mask.copy(), expand=expand)
            with np.errstate(all="ignore"):
                np.testing.assert_allclose(aicc, ST.AICc(ref[nc][0], 4 * nc, ref[nc][1]), rtol=1e-6, atol=1e-6,
                                           equal_nan=True)
    # planemask: only those pixels of the stored maps change
    uc._invalidate_stats()
    base = uc.get_AICc(2).copy()                             # default (model) mask
    pm = np.zeros((ny, nx), bool)
    pm[1, :] = True
    upd = uc.get_AICc(2, mask=mask.copy(), planemask=pm, expand=20)
    assert np.array_equal(upd[~pm], base[~pm], equal_nan=True)
    with np.errstate(all="ignore"):
        r2, n2 = ST.rss_maps(cube, m2.copy(), mask.copy(), expand=20)
        want = ST.AICc(r2[pm], 8, n2[pm])
    np.testing.assert_allclose(upd[pm], want, rtol=1e-6, atol=1e-6, equal_nan=True)


def test_two_cube_likelihood_with_zero_component_model(fitted_tile):
    """calc_AICc_likelihood(ucube, 1, 0, ucube_B=other) (UltraCube.py:1120-1216): the 0-component model is y = 0
    with p = 0 over the common mask (here: the 1-component model's own mask, expand 0)."""
    from mufasa_b200.cube import SpecCube
    from mufasa_b200.UltraCube import UltraCube, calc_AICc_likelihood
    t = fitted_tile
    cube = t["cube"]
    ucs = []
    for _ in range(2):
        uc = UltraCube(cube=SpecCube(cube, t["v"]), fittype=t["fittype"], device=0)
        p = uc.load_pcube()
        p.parcube, p.errcube = t["p1"].copy(), np.zeros_like(t["p1"])
        p.has_fit = np.any(p.parcube != 0, axis=0)
        p.fittype, p.npars = t["fittype"], 4
        uc.pcubes["1"] = p
        ucs.append(uc)
    got = calc_AICc_likelihood(ucs[0], 1, 0, ucube_B=ucs[1], expand=0)
    m1 = M.modelcube(t["v"], t["p1"], t["fittype"], dtype=np.float32)
    with np.errstate(all="ignore"):
        mask = m1 > 0
        r1, n1 = ST.rss_maps(cube, m1.copy(), mask.copy(), expand=0)
        r0, n0 = ST.rss_maps(cube, np.zeros(cube.shape), mask.copy(), expand=0)
        ref = ST.likelihood(ST.AICc(r1, 4, n1), ST.AICc(r0, 0, n0))
    ok = np.isfinite(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    # N can differ by the edge channels of the fp64 model mask: compare where the likelihood
    # is not sensitive to them, and exactly in sign
    assert np.nanmedian(np.abs(got[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))) < 1e-3
    assert np.array_equal(got[ok & (np.abs(ref) > 1)] > 0, ref[ok & (np.abs(ref) > 1)] > 0)
