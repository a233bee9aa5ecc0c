"""Convolved-cube stage without a GPU: the oracle (oracle/convolve.py) against independent scipy implementations and
the golden vectors generated from the reference's own sources (tests/golden/cnv_golden.npz), and the host logic of
mufasa_b200.convolve_tools / guess_refine / utils.fits_utils."""
import os

import numpy as np
import pytest
from scipy import ndimage as nd

from mufasa_b200 import guess_refine as gr
from mufasa_b200 import signals
from mufasa_b200.cube import SpecCube
from mufasa_b200.utils import fits_utils
from oracle import convolve as oc


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cnv_golden.npz"))


def test_downsample_header_matches_reference(gold):
    for crpix, cdelt, factor, p1, d1, p2, d2 in gold["downsample_header"]:
        h = {"CRPIX1": crpix, "CDELT1": cdelt, "CRPIX2": crpix + 3, "CDELT2": -cdelt}
        for impl in (oc.downsample_header, fits_utils.downsample_header):
            out = impl(impl(h, factor, 1), factor, 2)
            assert (out["CRPIX1"], out["CDELT1"], out["CRPIX2"], out["CDELT2"]) == (p1, d1, p2, d2)
        assert h["CRPIX1"] == crpix                       # the input header is not modified


def test_component_sorting_matches_reference(gold):
    data = gold["data"]
    np.testing.assert_array_equal(gr.mask_swap_2comp(data.copy(), gold["swapmask"]), gold["swapped"])
    for m in ("error_v", "Tpeak", "tautex", "tau", "chen2020"):
        np.testing.assert_array_equal(gr.quick_2comp_sort(data.copy(), filtsize=2, method=m), gold["sort_" + m], err_msg=m)
    np.testing.assert_array_equal(gr.refine_2c_guess(data[:8].copy(), f_sigv=0.5), gold["refine_2c"])


def test_pixel_map_of_a_downsampled_header():
    h = {"NAXIS1": 64, "NAXIS2": 48, "CRPIX1": 30.0, "CRPIX2": 20.0, "CDELT1": -0.002, "CDELT2": 0.002,
         "CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 52.0, "CRVAL2": 31.0}
    nh, coef = oc.downsample_map(h, 2)
    assert (nh["NAXIS1"], nh["NAXIS2"]) == (32, 24)
    (ay, by), (ax, bx) = fits_utils.linear_pixel_map(h, nh)
    assert (ay, by, ax, bx) == pytest.approx((2.0, 0.5, 2.0, 0.5), abs=1e-12)      # output pixel = 2 x 2 input pixels
    assert coef[1] == pytest.approx((ax, bx)) and coef[2] == pytest.approx((ay, by))
    grid = fits_utils.get_pixel_mapping(h, nh)
    assert grid.shape == (2, 24, 32) and grid[0, 3, 0] == pytest.approx(6.5) and grid[1, 0, 5] == pytest.approx(10.5)
    other = dict(nh, CRVAL1=53.0)
    with pytest.raises(NotImplementedError):
        fits_utils.linear_pixel_map(h, other)


def test_oracle_convolution_against_scipy():
    rng = np.random.default_rng(5)
    cube = rng.normal(size=(3, 21, 26))
    for kern in (oc.beam_kernel(3.0, 3.0, 0.0, 1.0, 2), oc.beam_kernel(4.0, 2.5, 30.0, 1.0, 2)):
        out = oc.convolve_planes(cube, kern)
        for c in range(cube.shape[0]):
            ref = nd.correlate(cube[c], kern / kern.sum(), mode='constant', cval=0.0)
            np.testing.assert_allclose(out[c], ref, rtol=1e-12, atol=1e-13)
    # NaNs are interpolated over with the kernel renormalised; the voxel itself stays NaN (the cube keeps its mask)
    cube[1, 5:8, 5:9] = np.nan
    kern = oc.beam_kernel(3.0, 3.0, 0.0, 1.0, 2)
    out = oc.convolve_planes(cube, kern)
    assert np.all(np.isnan(out[1, 5:8, 5:9])) and np.isfinite(out[1, 4, 5])
    y, x = 4, 6
    h = kern.shape[0] // 2
    pad = np.pad(cube[1], h, constant_values=0.0)
    win = pad[y:y + kern.shape[0], x:x + kern.shape[0]]
    ok = np.isfinite(win)
    assert out[1, y, x] == pytest.approx(np.sum(kern[ok] * win[ok]) / np.sum(kern[ok]), rel=1e-12)
    # a 2-D mask: excluded pixels act as NaN and come out NaN
    inc = np.ones(cube.shape[1:], bool)
    inc[:3] = False
    out2 = oc.convolve_planes(cube, kern, include=inc)
    assert np.all(np.isnan(out2[:, :3])) and np.all(np.isfinite(out2[0, 3:]))


def test_oracle_kernel_shape_and_orientation():
    k = oc.beam_kernel(4.0, 4.0, 0.0, 1.0, 2)
    sig = np.sqrt(3.0) * 4.0 / 2.3548200450309493
    assert k.shape[0] % 2 == 1 and k.shape[0] >= 8 * sig
    g = np.exp(-0.5 * (np.arange(k.shape[0]) - k.shape[0] // 2) ** 2 / sig ** 2)
    np.testing.assert_allclose(k / k.max(), np.outer(g, g), rtol=1e-12, atol=1e-300)
    # BPA = 0: the major axis points north (along y)
    e = oc.beam_kernel(6.0, 2.0, 0.0, 1.0, 2)
    c = e.shape[0] // 2
    assert e[c + 3, c] > e[c, c + 3]
    e90 = oc.beam_kernel(6.0, 2.0, 90.0, 1.0, 2)
    np.testing.assert_allclose(e90, e.T, rtol=1e-9, atol=1e-300)


def test_product_kernel_matches_oracle():
    from mufasa_b200.convolve_tools import Beam
    for bmaj, bmin, pa in ((3.2, 3.2, 0.0), (5.0, 3.0, 25.0), (4.0, 2.0, -70.0)):
        ok = oc.beam_kernel(bmaj, bmin, pa, 0.8, 2)
        kern = Beam(2 * bmaj, 2 * bmin, pa).deconvolve(Beam(bmaj, bmin, pa)).as_kernel(0.8)
        k2 = kern["k2"] if "k2" in kern else np.outer(kern["ky"], kern["kx"])
        assert k2.shape == ok.shape
        np.testing.assert_allclose(k2 / k2.sum(), ok / ok.sum(), rtol=2e-6, atol=1e-12)
    with pytest.raises(ValueError):
        Beam(3.0, 3.0).deconvolve(Beam(3.0, 3.0))


def test_oracle_bilinear_against_scipy():
    rng = np.random.default_rng(6)
    cube = rng.normal(size=(2, 17, 23))
    out = oc.resample_bilinear(cube, 8, 11, 2.0, 0.5, 2.0, 0.5)
    yy, xx = np.meshgrid(2.0 * np.arange(8) + 0.5, 2.0 * np.arange(11) + 0.5, indexing="ij")
    for c in range(2):
        ref = nd.map_coordinates(cube[c], [yy, xx], order=1, mode='nearest')
        np.testing.assert_allclose(out[c], ref, rtol=1e-12, atol=1e-13)
    # factor 2: every output pixel is the mean of a 2 x 2 block
    np.testing.assert_allclose(out[0, 3, 4], cube[0, 6:8, 8:10].mean(), rtol=1e-12)
    # beyond half a pixel outside -> NaN; a NaN neighbour -> NaN
    far = oc.resample_bilinear(cube, 10, 11, 2.0, 0.5, 2.0, 0.5)
    assert np.all(np.isnan(far[:, 9])) and np.all(np.isfinite(far[:, 8, :11]))     # y = 18.5 > 16.5; y = 16.5 is the edge
    cube[0, 6, 8] = np.nan
    assert np.isnan(oc.resample_bilinear(cube, 8, 11, 2.0, 0.5, 2.0, 0.5)[0, 3, 4])


def test_small_object_morphology():
    m = np.zeros((12, 12), bool)
    m[2:9, 2:9] = True
    m[4, 4] = False
    m[0, 0] = True
    m[10, 10] = m[11, 11] = True                     # diagonal neighbours are separate objects (4-connectivity)
    out = signals.remove_small_objects(m, 9)
    assert not out[0, 0] and not out[10, 10] and out[3, 3] and not out[4, 4]
    filled = signals.remove_small_holes(m, 9)
    assert filled[4, 4] and not filled[0, 5]
    np.testing.assert_array_equal(signals.remove_small_objects(np.zeros((4, 4), bool), 3), np.zeros((4, 4), bool))


def test_speccube_mask_view_and_fits_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    data = rng.normal(size=(5, 6, 7)).astype(np.float32)
    data[:, 0, 0] = np.nan
    v = np.linspace(-1.0, 1.0, 5)
    hdr = {"CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 10.0, "CRVAL2": 20.0, "CRPIX1": 3.0, "CRPIX2": 4.0,
           "CDELT1": -0.002, "CDELT2": 0.002, "CUNIT1": "deg", "CUNIT2": "deg", "BMAJ": 0.008, "BMIN": 0.008, "BPA": 0.0}
    cube = SpecCube(data, v, rest_value=23.6944955e9, header=hdr)
    plane = np.ones((6, 7), bool)
    plane[:, -1] = False
    view = cube.with_mask(plane)
    assert view._data is cube._data and cube.plane_include is None
    inc = view.mask.include()
    assert not inc[:, 0, 0].any() and not inc[:, :, -1].any() and inc[:, 2, 2].all()
    path = str(tmp_path / "cube.fits")
    cube.write(path)
    back = SpecCube.read(path)
    np.testing.assert_array_equal(back._data, data)
    np.testing.assert_allclose(back.spectral_axis, v, rtol=0, atol=1e-12)
    assert back.rest_value == 23.6944955e9 and back.header["BMAJ"] == 0.008 and back.header["CTYPE1"] == "RA---TAN"


def test_guess_from_cnvpara_fills_the_target_grid():
    rng = np.random.default_rng(8)
    ny, nx = 16, 18
    par = np.zeros((8, ny, nx))
    yy, xx = np.mgrid[:ny, :nx]
    blob = (yy - 8) ** 2 + (xx - 9) ** 2 < 36
    par[0] = 0.05 * xx - 0.3
    par[1] = 0.3
    par[2] = 6.0
    par[3] = 2.0
    par[4:] = rng.uniform(0.01, 0.02, (4, ny, nx))
    par[:, ~blob] = 0.0
    par[:, 8, 9] = 0.0                                   # a hole inside the footprint
    h = {"NAXIS1": 2 * nx, "NAXIS2": 2 * ny, "CRPIX1": 18.0, "CRPIX2": 16.0, "CDELT1": -0.002, "CDELT2": 0.002,
         "CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 52.0, "CRVAL2": 31.0}
    hc, _ = oc.downsample_map(h, 2)
    mask = signals.remove_small_holes(np.any(par != 0, axis=0))
    g = gr.guess_from_cnvpara(par, hc, h, clean_map=False, tau_thresh=1, mask=mask)
    assert g.shape == (4, 2 * ny, 2 * nx)
    fine = np.all(np.isfinite(g), axis=0)
    big = np.kron(blob, np.ones((2, 2), bool))
    assert fine[big].mean() > 0.95 and fine[16:18, 18:20].all()      # the hole was interpolated over
    # the velocity gradient survives the regridding: 0.05 km/s per coarse pixel = 0.025 per fine pixel
    row = g[0, 16, 10:26]
    assert np.allclose(np.diff(row), 0.025, atol=5e-3)
    assert np.nanmax(np.abs(g[1][fine] - 0.3)) < 1e-6 and np.nanmax(np.abs(g[2][fine] - 6.0)) < 1e-6


def test_regrid_many_equals_map_by_map_regrid():
    from mufasa_b200.convolve_tools import regrid, regrid_many
    rng = np.random.default_rng(9)
    ny, nx = 20, 24
    maps = rng.normal(size=(5, ny, nx))
    yy, xx = np.mgrid[:ny, :nx]
    maps[:, (yy - 10) ** 2 + (xx - 12) ** 2 > 81] = np.nan
    maps[3:, 5, 5] = np.nan                              # two maps with a different footprint
    h = {"NAXIS1": 2 * nx, "NAXIS2": 2 * ny, "CRPIX1": 24.0, "CRPIX2": 20.0, "CDELT1": -0.002, "CDELT2": 0.002,
         "CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 52.0, "CRVAL2": 31.0}
    hc, _ = oc.downsample_map(h, 2)
    many = regrid_many(maps, hc, h)
    for k in range(5):
        np.testing.assert_array_equal(many[k], regrid(maps[k], hc, h))


def test_beam_deconvolution_at_any_position_angle():
    """Beam.deconvolve (radio_beam's relations for two elliptical Gaussians, convolve_tools.py:90-138) undoes
    Beam.convolve (second-moment matrices add -- an independent route) for beams at different position angles."""
    from mufasa_b200.convolve_tools import Beam
    rng = np.random.default_rng(1)
    for _ in range(200):
        maj = rng.uniform(1, 3)
        b2 = Beam(maj, maj * rng.uniform(0.3, 1), rng.uniform(-90, 90))
        majk = rng.uniform(1, 3)
        k = Beam(majk, majk * rng.uniform(0.3, 1), rng.uniform(-90, 90))
        d = b2.convolve(k).deconvolve(b2)
        dpa = abs(((d.pa - k.pa) + 90) % 180 - 90)
        assert abs(d.major - k.major) < 1e-9 and abs(d.minor - k.minor) < 1e-9
        assert np.radians(dpa) * (k.major - k.minor) < 1e-9          # the angle matters only for elongated beams
    same = Beam(5, 3, 30).deconvolve(Beam(3, 2, 30))
    assert abs(same.major - 4.0) < 1e-12 and abs(same.minor - np.sqrt(5.0)) < 1e-12 and same.pa == 30.0
    with pytest.raises(ValueError):
        Beam(3, 2, 10).deconvolve(Beam(3, 2.5, 60))
