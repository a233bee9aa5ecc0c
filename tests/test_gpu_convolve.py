"""GPU parity of the convolved-cube stage through the C ABI: b200fit_convolve_planes and
b200fit_resample_bilinear against oracle/convolve.py (float64 numpy), the composed convolve_sky_byfactor, and the
two-step fit of iter_2comp_fit (fit the convolved cube, turn its parameters into guesses, fit at full resolution).

Tolerance: the kernels accumulate in fp32 over K^2 <= 4225 terms of the fp32 cube, the oracle in fp64 --
|gpu - oracle| <= 2e-6 * max|cube| + 2e-6 |oracle| is asserted; NaN patterns must agree exactly."""
import os

import numpy as np
import pytest

import _tiles as T
from mufasa_b200 import convolve_tools as ct
from mufasa_b200 import guess_refine as gr
from mufasa_b200 import synth
from mufasa_b200.cube import SpecCube
from mufasa_b200.UltraCube import UltraCube
from oracle import convolve as oc

pytestmark = pytest.mark.gpu

HDR = {"CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 52.0, "CRVAL2": 31.0, "CUNIT1": "deg", "CUNIT2": "deg",
       "CDELT1": -0.0024, "CDELT2": 0.0024, "BMAJ": 0.0088, "BMIN": 0.0088, "BPA": 0.0}


@pytest.fixture(scope="module")
def engine():
    import torch
    from mufasa_b200.engine import get_engine
    assert torch.cuda.is_available()
    return get_engine(0)


def _close(gpu, ref, scale):
    assert np.array_equal(np.isnan(gpu), np.isnan(ref))
    ok = np.isfinite(ref)
    assert np.all(np.abs(gpu[ok] - ref[ok]) <= 2e-6 * scale + 2e-6 * np.abs(ref[ok]))


def _header(ny, nx, **kw):
    return dict(HDR, NAXIS1=nx, NAXIS2=ny, CRPIX1=nx / 2.0, CRPIX2=ny / 2.0, **kw)


@pytest.mark.parametrize("nchan,ny,nx,beam", [(5, 37, 45, (3.0, 3.0, 0.0)), (3, 64, 32, (1.0, 1.0, 0.0)),
                                              (4, 33, 70, (5.0, 3.0, 30.0)), (2, 20, 19, (10.0, 10.0, 0.0)),
                                              (2, 40, 41, (10.0, 5.0, -50.0))])
def test_convolve_planes_matches_oracle(engine, nchan, ny, nx, beam):
    import torch
    rng = np.random.default_rng(ny * nx)
    cube = rng.normal(size=(nchan, ny, nx)).astype(np.float32)
    cube[rng.random(cube.shape) < 0.03] = np.nan                  # scattered bad voxels
    cube[:, : ny // 5, : nx // 4] = np.nan                        # a blank corner, as in a mosaic
    inc = np.ones((ny, nx), bool)
    inc[-3:] = False
    bmaj, bmin, pa = beam
    k64 = oc.beam_kernel(bmaj, bmin, pa, 1.0, 2)
    kern = ct.Beam(2 * bmaj, 2 * bmin, pa).deconvolve(ct.Beam(bmaj, bmin, pa)).as_kernel(1.0)
    assert ("kx" in kern) == (bmaj == bmin)                       # circular beam -> separable path
    kdev = {k: torch.from_numpy(v).cuda() for k, v in kern.items()}
    d = torch.from_numpy(cube).cuda()
    for include in (None, inc):
        dinc = None if include is None else torch.from_numpy(include.astype(np.uint8)).cuda()
        out = engine.convolve_planes(d, include=dinc, **kdev).cpu().numpy()
        _close(out, oc.convolve_planes(cube, k64, include=include), np.nanmax(np.abs(cube)))


@pytest.mark.parametrize("K,nchan,ny,nx", [(9, 3, 64, 64), (11, 2, 70, 132), (13, 2, 129, 68), (15, 3, 40, 200),
                                            (23, 4, 150, 136), (59, 2, 96, 72), (65, 1, 130, 128)])
def test_convolve_fast_separable_path_matches_oracle(engine, K, nchan, ny, nx):
    """Planes with 16-byte aligned rows and no pixel mask take the fast separable kernel (cp.async window, packed
    FMAs on pairs, static ramps for each K mod 8) with the generic kernel redoing the tiles that hold NaNs: every
    residue of K mod 8, tiles cut by the image edge, all-finite planes, scattered NaNs and a blank corner; and the
    result equals the generic kernel's to rounding."""
    import os
    import torch
    rng = np.random.default_rng(K * 1000 + nx)
    h = K // 2
    kx = np.exp(-0.5 * ((np.arange(K) - h) / (K / 6.0)) ** 2)
    ky = np.exp(-0.5 * ((np.arange(K) - h) / (K / 7.0)) ** 2)      # different widths: a transposed pass would show
    k64 = np.outer(ky, kx)
    kdev = dict(kx=torch.from_numpy(kx.astype(np.float32)).cuda(), ky=torch.from_numpy(ky.astype(np.float32)).cuda())
    k64 = np.outer(ky.astype(np.float32).astype(np.float64), kx.astype(np.float32).astype(np.float64))
    clean = rng.normal(size=(nchan, ny, nx)).astype(np.float32)
    holes = clean.copy()
    holes[rng.random(holes.shape) < 0.002] = np.nan
    holes[:, : ny // 5, : nx // 4] = np.nan
    holes[-1, -1, -1] = np.nan                                     # a NaN in the last corner of the last plane
    for cube in (clean, holes):
        d = torch.from_numpy(cube).cuda()
        out = engine.convolve_planes(d, **kdev).cpu().numpy()
        _close(out, oc.convolve_planes(cube, k64), np.nanmax(np.abs(cube)))
        os.environ["B200FIT_CONV_GENERIC"] = "1"
        try:
            gen = engine.convolve_planes(d, **kdev).cpu().numpy()
        finally:
            del os.environ["B200FIT_CONV_GENERIC"]
        assert np.array_equal(np.isnan(out), np.isnan(gen))
        assert np.nanmax(np.abs(out - gen)) <= 1e-5


def test_convolve_separable_equals_general_path(engine):
    import torch
    rng = np.random.default_rng(3)
    cube = rng.normal(size=(6, 50, 47)).astype(np.float32)
    cube[rng.random(cube.shape) < 0.05] = np.nan
    kern = ct.Beam(8.0, 8.0).deconvolve(ct.Beam(4.0, 4.0)).as_kernel(1.0)
    d = torch.from_numpy(cube).cuda()
    sep = engine.convolve_planes(d, kx=torch.from_numpy(kern["kx"]).cuda(), ky=torch.from_numpy(kern["ky"]).cuda())
    k2 = np.outer(kern["ky"], kern["kx"]).astype(np.float32)
    full = engine.convolve_planes(d, k2=torch.from_numpy(k2).cuda())
    _close(sep.cpu().numpy(), full.cpu().numpy().astype(np.float64), 5.0)


def test_convolve_properties_at_cube_scale(engine):
    """Size-independent properties on a 256 x 256 x 64 cube: a constant cube stays constant away from the border
    (the kernel is normalised), next to the border it is dimmed by the zero fill, and the operator is linear."""
    import torch
    ny = nx = 256
    nchan = 64
    kern = ct.Beam(2 * 3.6, 2 * 3.6).deconvolve(ct.Beam(3.6, 3.6)).as_kernel(1.0)
    K = kern["kx"].size
    kx, ky = (torch.from_numpy(kern[k]).cuda() for k in ("kx", "ky"))
    ones = torch.full((nchan, ny, nx), 2.5, dtype=torch.float32, device="cuda")
    out = engine.convolve_planes(ones, kx=kx, ky=ky)
    h = K // 2
    assert torch.allclose(out[:, h:-h, h:-h], ones[:, h:-h, h:-h], rtol=0, atol=2e-6)
    inside = float(kern["ky"][h:].sum() / kern["ky"].sum())        # rows 0..h of the window lie in the image
    assert float(out[0, 0, nx // 2]) == pytest.approx(2.5 * inside, rel=1e-5)
    assert float(out[0, 0, 0]) == pytest.approx(2.5 * inside * inside, rel=1e-5)
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn((nchan, ny, nx), device="cuda", generator=g)
    b = torch.randn((nchan, ny, nx), device="cuda", generator=g)
    lhs = engine.convolve_planes((2 * a - 3 * b).contiguous(), kx=kx, ky=ky)
    rhs = 2 * engine.convolve_planes(a, kx=kx, ky=ky) - 3 * engine.convolve_planes(b, kx=kx, ky=ky)
    assert float((lhs - rhs).abs().max()) < 2e-5


@pytest.mark.parametrize("ny,nx,factor", [(37, 45, 2), (64, 64, 2), (50, 41, 3)])
def test_resample_bilinear_matches_oracle(engine, ny, nx, factor):
    import torch
    rng = np.random.default_rng(ny)
    cube = rng.normal(size=(4, ny, nx)).astype(np.float32)
    cube[1, ny // 2, nx // 2] = np.nan
    nh, coef = oc.downsample_map(_header(ny, nx), factor)
    (ax, bx), (ay, by) = coef[1], coef[2]
    out = engine.resample_bilinear(torch.from_numpy(cube).cuda(), nh["NAXIS2"], nh["NAXIS1"], ay, by, ax, bx)
    ref = oc.resample_bilinear(cube, nh["NAXIS2"], nh["NAXIS1"], ay, by, ax, bx)
    _close(out.cpu().numpy(), ref, np.nanmax(np.abs(cube)))


def test_convolve_sky_byfactor_matches_oracle_composition(engine, tmp_path):
    """The whole stage on a 48 x 52 x 40 cube with a blank border: edge trim (disk(5) erosion of the footprint),
    smoothing to 2 x the beam, 2 x downsampling, header bookkeeping, FITS output."""
    from mufasa_b200 import signals
    rng = np.random.default_rng(12)
    nchan, ny, nx = 40, 48, 52
    cube = rng.normal(size=(nchan, ny, nx)).astype(np.float32)
    cube[:, :4] = np.nan
    cube[:, :, -6:] = np.nan
    hdr = _header(ny, nx)
    sc = SpecCube(cube, np.linspace(-2, 2, nchan), rest_value=23.6944955e9, header=hdr)
    path = str(tmp_path / "cnv.fits")
    new = ct.convolve_sky_byfactor(sc, 2, savename=path, edgetrim_width=5)
    foot = signals.binary_erosion(np.isfinite(cube).any(axis=0), signals.disk(5))
    k64 = oc.beam_kernel(hdr["BMAJ"], hdr["BMIN"], 0.0, hdr["CDELT2"], 2)
    ref = oc.convolve_planes(cube, k64, include=foot)
    nh, coef = oc.downsample_map(hdr, 2)
    ref = oc.resample_bilinear(ref, nh["NAXIS2"], nh["NAXIS1"], coef[2][0], coef[2][1], coef[1][0], coef[1][1])
    assert new.shape == (nchan, 24, 26)
    _close(new._data, ref, np.nanmax(np.abs(cube)))
    assert new.header["CDELT2"] == pytest.approx(0.0048) and new.header["CRPIX1"] == nh["CRPIX1"]
    assert new.header["BMAJ"] == pytest.approx(2 * hdr["BMAJ"])
    back = SpecCube.read(path)
    np.testing.assert_array_equal(back._data, new._data)
    np.testing.assert_allclose(back.spectral_axis, sc.spectral_axis, atol=1e-12)
    # no downsampling: same grid, same header geometry
    same = ct.convolve_sky_byfactor(sc, 2, edgetrim_width=None, downsample=False)
    _close(same._data, oc.convolve_planes(cube, k64), np.nanmax(np.abs(cube)))


def test_snr_mask_matches_numpy_mirror(engine):
    """snr_mask (convolve_tools.py:141-178): the GPU supplies mad_std and the peak per pixel; compare with numpy."""
    from mufasa_b200 import signals
    from mufasa_b200.utils.convolve import convolve_gauss2d
    syn = synth.make_cube(3, T.oracle_modelcube, shape=(40, 44, 300))
    sc = SpecCube(syn["cube"], syn["v"], header=_header(40, 44))
    got = ct.snr_mask(sc, snr_min=3.0)
    d = syn["cube"].astype(np.float64)
    err = signals.mad_std(d, axis=0)
    peaksnr = convolve_gauss2d(np.nanmax(d / err, axis=0), 1)
    pm = signals.binary_erosion(peaksnr > 3.0, signals.disk(1))
    im = signals.remove_small_objects(pm, 9)
    pm = im if im.sum() > 9 else pm
    ref = signals.binary_dilation(pm, signals.disk(3))
    assert np.array_equal(got, ref) and 0 < ref.sum()


def test_two_step_fit_through_the_convolved_cube(engine):
    """iter_2comp_fit's recipe (master_fitter.py:682-775) for one component: fit the 2 x convolved cube from moment
    guesses, turn its parameter maps into full-resolution guesses (guess_from_cnvpara), fit the original cube with
    them.  The guesses must land near the truth and the final fit must recover it."""
    ny, nx, nchan = 48, 48, 500
    syn = synth.make_cube(1, T.oracle_modelcube, shape=(ny, nx, nchan))
    hdr = _header(ny, nx)
    sc = SpecCube(syn["cube"], syn["v"], header=hdr)
    uc = UltraCube(cube=sc, fittype=syn["fittype"], device=0)
    uc.convolve_cube(factor=2, edgetrim_width=5)
    assert uc.cube_cnv.shape == (nchan, ny // 2, nx // 2)
    ucc = UltraCube(cube=uc.cube_cnv, fittype=syn["fittype"], device=0)
    ucc.fit_cube(ncomp=1, snr_min=3.0)
    pc = ucc.pcubes["1"]
    assert pc.has_fit.sum() > 100
    para_cnv = np.append(pc.parcube, pc.errcube, axis=0)
    from mufasa_b200 import signals
    mask = signals.remove_small_holes(np.any(np.isfinite(para_cnv) & (para_cnv != 0), axis=0))
    para_cnv[para_cnv == 0] = np.nan
    guesses = gr.guess_from_cnvpara(para_cnv, uc.cube_cnv.header, sc.header, clean_map=True, tau_thresh=1, mask=mask)
    assert guesses.shape == (4, ny, nx)
    fin = np.all(np.isfinite(guesses), axis=0)
    assert fin.sum() > 0.4 * ny * nx
    truth = syn["truth1"]
    # the 48-pixel truth field swings by 0.4 km/s over ~20 pixels; smoothing it with the 2 x beam and the guess
    # refinement leaves ~0.1 km/s of error, well inside the basin of the fit (line widths 0.12-0.6 km/s)
    assert np.median(np.abs(guesses[0][fin] - truth[0][fin])) < 0.25
    uc.fit_cube(ncomp=1, guesses=guesses, snr_min=0.0)
    p = uc.pcubes["1"]
    good = p.has_fit & fin
    assert good.sum() > 0.4 * ny * nx
    dv = np.abs(p.parcube[0][good] - truth[0][good])
    assert np.median(dv) < 0.02 and np.median(np.abs(p.parcube[1][good] - truth[1][good])) < 0.03


def test_region_iter_2comp_fit_from_fits_files(engine, tmp_path):
    """Region + iter_2comp_fit (master_fitter.py:65-126, :682-775) from a FITS cube on disk: convolved cube and its
    1- / 2-component parameter maps are written next to it, the full-resolution fits start from guesses regridded
    from them, everything is saved in the reference's ``{root}_{n}vcomp.fits`` layout and can be resumed."""
    import os
    from mufasa_b200 import master_fitter as mf
    from mufasa_b200.utils import fits_io
    ny, nx, nchan = 48, 48, 500
    syn = synth.make_cube(3, T.oracle_modelcube, shape=(ny, nx, nchan))
    sc = SpecCube(syn["cube"], syn["v"], rest_value=23.6944955e9, header=_header(ny, nx))
    path = str(tmp_path / "field.fits")
    sc.write(path)
    reg = mf.Region(path, "field", paraDir=str(tmp_path / "paras"), fittype=syn["fittype"], device=0)
    mf.iter_2comp_fit(reg, snr_min=3.0)
    for f in ("field_conv2Xbeam.fits", "paras/field_conv2Xbeam_1vcomp.fits", "paras/field_conv2Xbeam_2vcomp.fits",
              "paras/field_1vcomp.fits", "paras/field_2vcomp.fits"):
        assert os.path.isfile(str(tmp_path / f)), f
    data, hdr = fits_io.read(str(tmp_path / "paras/field_2vcomp.fits"))
    assert data.shape == (16, ny, nx) and hdr.value("PLANE1") is not None
    assert reg.ucube_cnv.cube.shape == (nchan, ny // 2, nx // 2)
    p1, p2 = reg.ucube.pcubes["1"], reg.ucube.pcubes["2"]
    assert p1.has_fit.sum() > 0.3 * ny * nx and p2.has_fit.sum() > 0.2 * ny * nx
    one = p1.has_fit & (syn["ncomp_true"] == 1)
    assert one.sum() > 100
    assert np.median(np.abs(p1.parcube[0][one] - syn["truth1"][0][one])) < 0.03
    # model selection on top of the two-step fits picks 2 components mostly where the truth has 2
    reg.ucube.get_best_2c_parcube()
    two_true = (syn["ncomp_true"] == 2) & p2.has_fit
    if two_true.sum() > 50:
        assert (reg.ucube.ncomp_map[two_true] == 2).mean() > (reg.ucube.ncomp_map[one] == 2).mean()
    # resume: a second Region reads the saved maps instead of fitting
    reg2 = mf.Region(path, "field", paraDir=str(tmp_path / "paras"), fittype=syn["fittype"], device=0)
    reg2.ucube.get_model_fit([1, 2], update=False)
    np.testing.assert_array_equal(reg2.ucube.pcubes["2"].parcube, p2.parcube)


def test_master_2comp_fit_pipeline(engine, tmp_path):
    """The full driver (master_fitter.py:583-679): two-step fit, bad-pixel and marginal refits,
    expansion into the surroundings, products on disk."""
    import os
    from mufasa_b200 import master_fitter as mf
    from mufasa_b200.utils import fits_io
    ny, nx, nchan = 48, 48, 500
    syn = synth.make_cube(3, T.oracle_modelcube, shape=(ny, nx, nchan))
    path = str(tmp_path / "field.fits")
    SpecCube(syn["cube"], syn["v"], rest_value=23.6944955e9, header=_header(ny, nx)).write(path)
    reg = mf.Region(path, "field", paraDir=str(tmp_path / "paras"), fittype=syn["fittype"], device=0)
    mf.iter_2comp_fit(reg, snr_min=3.0)
    n1_before = int(reg.ucube.pcubes["1"].has_fit.sum())
    lnk21_before = np.array(reg.ucube.get_AICc_likelihood(2, 1))
    # the remaining stages, as master_2comp_fit strings them together
    mf.refit_bad_2comp(reg, snr_min=1.0, lnk_thresh=-5)
    for nc in (1, 2):
        mf.refit_marginal(reg, ncomp=nc, lnk_thresh=10, holes_only=False, method='best_neighbour')
    mf.fit_surroundings(reg, ncomps=[1, 2], snr_min=3.0, max_iter=3, match_footprint=True)
    mf.save_best_2comp_fit(reg)
    p1, p2 = reg.ucube.pcubes["1"], reg.ucube.pcubes["2"]
    assert int(p1.has_fit.sum()) >= n1_before                      # refits only ever replace or add fits
    lnk21_after = reg.ucube.get_AICc_likelihood(2, 1)
    both = np.isfinite(lnk21_before) & np.isfinite(lnk21_after)
    assert np.nanmedian(lnk21_after[both] - lnk21_before[both]) >= -1e-6
    for key in ("2vcomp_final", "lnk21", "lnk10", "lnk20", "SNR", "model_mom0", "mom0", "chi2red_1c", "chi2red_2c"):
        f = str(tmp_path / "paras" / ("field_%s.fits" % key))
        assert os.path.isfile(f), key
    final, hdr = fits_io.read(str(tmp_path / "paras" / "field_2vcomp_final.fits"))
    assert final.shape == (16, ny, nx)
    truth = syn["ncomp_true"]
    two = np.isfinite(final[4])                                   # second component kept by the selection
    fitted = p1.has_fit
    assert (two[fitted] == (truth[fitted] == 2)).mean() > 0.85
    snr, _ = fits_io.read(str(tmp_path / "paras" / "field_SNR.fits"))
    assert np.nanmax(snr) > 10 and snr.shape == (ny, nx)
    # the device-side product maps against the reference's host recipe on full model cubes
    from mufasa_b200.UltraCube import get_masked_moment
    modbest = mf.get_best_2comp_model(reg)
    snr_host = mf.get_best_2comp_snr_mod(reg, modbest)
    sel = reg.ucube.ncomp_map > 0
    np.testing.assert_allclose(snr[sel], snr_host[sel], rtol=1e-5)
    mm0, _ = fits_io.read(str(tmp_path / "paras" / "field_model_mom0.fits"))
    dv = syn["v"][1] - syn["v"][0]
    np.testing.assert_allclose(mm0, np.nansum(modbest.astype(np.float64), axis=0) * dv, rtol=1e-5, atol=1e-6)
    m0, _ = fits_io.read(str(tmp_path / "paras" / "field_mom0.fits"))
    m0_host = get_masked_moment(reg.ucube.cube, modbest.astype(np.float64), order=0, expand=20)
    assert np.array_equal(np.isnan(m0), np.isnan(m0_host))
    np.testing.assert_allclose(m0[np.isfinite(m0)], m0_host[np.isfinite(m0)], rtol=1e-6, atol=1e-5)
    # and the one-call driver runs the same chain
    reg2 = mf.Region(path, "field2", paraDir=str(tmp_path / "paras2"), fittype=syn["fittype"], device=0)
    mf.master_2comp_fit(reg2, snr_min=3.0, max_expand_iter=2)
    assert os.path.isfile(str(tmp_path / "paras2" / "field2_2vcomp_final.fits"))


def _wide_field(tmp_path, ny=40, nx=40, nchan=500):
    """A field whose centre holds a second, fainter component 3 km/s from the first; everywhere the 2-component
    fit is started from two guesses on top of the FIRST component, so it cannot find the second one."""
    from mufasa_b200 import master_fitter as mf
    v = synth.vaxis(nchan, 0.0724)
    yy, xx = np.mgrid[0:ny, 0:nx]
    par1 = np.stack([0.3 * (xx / nx - 0.5) + 0.013, np.full((ny, nx), 0.3), np.full((ny, nx), 7.0), np.full((ny, nx), 3.0)])
    par2 = np.stack([par1[0] + 3.0, np.full((ny, nx), 0.25), np.full((ny, nx), 6.0), np.full((ny, nx), 1.5)])
    wide = np.hypot(yy - ny / 2, xx - nx / 2) < 12
    m1 = T.oracle_modelcube(par1, v, "nh3_multi_v")
    m2 = T.oracle_modelcube(np.concatenate([par1, par2]), v, "nh3_multi_v")
    rng = np.random.default_rng(11)
    cube = (np.where(wide[None], m2, m1) + rng.normal(0, 0.1, size=m1.shape)).astype(np.float32)
    path = str(tmp_path / "wide.fits")
    SpecCube(cube, v, rest_value=23.6944955e9, header=_header(ny, nx)).write(path)
    reg = mf.Region(path, "wide", paraDir=str(tmp_path / "paras"), fittype="nh3_multi_v", device=0)
    g1 = par1.copy()
    g2 = np.concatenate([par1, par1])
    g2[0] -= 0.15
    g2[4] += 0.15
    g2[3] *= 0.5
    g2[7] *= 0.5
    reg.ucube.fit_cube(ncomp=[1], simpfit=True, guesses=g1, snr_min=0.0)
    reg.ucube.fit_cube(ncomp=[2], simpfit=True, guesses=g2, snr_min=0.0)
    for nc in (1, 2):
        reg.ucube.paraPaths[str(nc)] = '{}/{}_{}vcomp.fits'.format(reg.ucube.paraDir, reg.ucube.paraNameRoot, nc)
    return reg, wide, par1, par2


def test_refit_2comp_wide_recovers_a_distant_second_component(engine, tmp_path):
    """refit_2comp_wide (master_fitter.py:981-1098), both methods.  'moments': guesses from the spectrum with its
    main component blanked.  'residual': the 1-component fit of the convolved residual of the best model, regridded to
    full resolution -- with the residual mask as the reference's docstrings describe it, and with the reference's
    actual branches, under which a residual with significant pixels ends the stage without a refit."""
    from mufasa_b200 import master_fitter as mf
    reg, wide, par1, par2 = _wide_field(tmp_path)
    p2 = reg.ucube.pcubes['2']
    sep0 = np.abs(p2.parcube[4] - p2.parcube[0])
    assert np.median(sep0[wide]) < 1.5                          # the badly started fit did not find the far line
    lnk21_0 = np.array(reg.ucube.get_AICc_likelihood(2, 1))
    # the reference's control flow: the residual cube has significant pixels -> returned without a mask -> SNRMaskError
    assert mf.RESIDUAL_MASK_AS_DOCUMENTED is False
    assert mf.refit_2comp_wide(reg, snr_min=3, method='residual', planemask=wide) is None
    assert reg.progress_log[-1]["n_attempted"] == 0 and not hasattr(reg, 'ucube_res_cnv')
    np.testing.assert_array_equal(np.abs(p2.parcube[4] - p2.parcube[0]), sep0)
    # as documented: masked to the significant residual, convolved, fitted, regridded, refitted
    mf.RESIDUAL_MASK_AS_DOCUMENTED = True
    try:
        good = mf.refit_2comp_wide(reg, snr_min=3, method='residual', planemask=wide)
    finally:
        mf.RESIDUAL_MASK_AS_DOCUMENTED = False
    assert good is not None and good.shape == wide.shape and good[wide].mean() > 0.8 and not good[~wide].any()
    assert reg.ucube_res_cnv.cube.shape[1:] == (wide.shape[0] // 2, wide.shape[1] // 2)
    v_far = np.maximum(p2.parcube[0], p2.parcube[4])
    assert np.median(np.abs(v_far[good] - par2[0][good])) < 0.05
    lnk21_1 = reg.ucube.get_AICc_likelihood(2, 1)
    assert np.all(lnk21_1[good] > lnk21_0[good]) and np.median(lnk21_1[good]) > 5
    np.testing.assert_array_equal(lnk21_1[~wide & np.isfinite(lnk21_0)], lnk21_0[~wide & np.isfinite(lnk21_0)])
    assert os.path.isfile(str(tmp_path / "paras" / "wide_1vcomp_onBestResidual.fits"))
    assert os.path.isfile(str(tmp_path / "paras" / "wide_2vcomp.fits"))
    # 'moments' on a fresh copy of the same field
    reg_m, wide, par1, par2 = _wide_field(tmp_path)
    good_m = mf.refit_2comp_wide(reg_m, snr_min=3, method='moments', planemask=wide, save_para=False)
    p2m = reg_m.ucube.pcubes['2']
    assert good_m is not None and good_m[wide].mean() > 0.5 and not good_m[~wide].any()
    assert np.median(np.abs(np.maximum(p2m.parcube[0], p2m.parcube[4])[good_m] - par2[0][good_m])) < 0.05
    with pytest.raises(Exception):
        mf.refit_2comp_wide(reg_m, method='nonsense', planemask=wide)
