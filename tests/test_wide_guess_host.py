"""Host logic of the wide-separation recovery (mufasa/moment_guess.py:119-367, 523-701): the windowed moments in
pyspeckit's and spectral_cube's definitions against the oracle's per-spectrum loops (oracle/guesses.py; they restate
those packages' definitions [RECALLED -- pyspeckit / spectral_cube are absent: parity unpinned]), moment_guesses_1c against the golden vectors made by
executing the reference, and mom_guess_wide_sep on a spectrum with two well separated lines."""
import os

import numpy as np

from mufasa_b200 import moment_guess as mmg
from oracle import guesses as OG
from mufasa_b200.cube import SpecCube

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pys_moments_loop(d, v):
    return OG.pyspeckit_moments(v, d)


def _cube(seed=0, nchan=120, ny=5, nx=6):
    rng = np.random.default_rng(seed)
    v = (np.arange(nchan) - nchan // 2) * 0.1
    cen = rng.uniform(-2, 2, size=(ny, nx))
    data = 2.0 * np.exp(-0.5 * ((v[:, None, None] - cen) / 0.3) ** 2) + 0.05 * rng.normal(size=(nchan, ny, nx))
    data[rng.random(data.shape) < 0.02] = np.nan
    data[:, 0, 0] = np.nan
    return SpecCube(data.astype(np.float32), v), cen


def test_window_moments_match_the_pyspeckit_definition():
    cube, cen = _cube()
    vmap = cen + 0.2
    vmap[1, 1] = np.nan                                        # no window -> pixel keeps momentcube's zeros
    m0, m1, m2 = mmg.window_moments(cube, window_hwidth=1.5, v_atpeak=vmap)
    v = cube.spectral_axis
    ref = OG.window_moments_map(cube._data, v, vmap, 1.5)
    np.testing.assert_allclose(np.array([m0, m1, m2]), ref, rtol=1e-10, atol=1e-12)
    assert m0[0, 0] == 0 and m0[1, 1] == 0
    good = np.isfinite(vmap) & (m0 != 0)
    assert np.abs(m1[good] - cen[good]).max() < 0.1            # the centroid finds the line ...
    assert np.median(np.abs(m2[good] - 0.3)) < 0.05             # ... and the width its sigma, roughly
    # without a velocity map the whole-spectrum centroid is the window centre
    a0, a1, a2 = mmg.window_moments(cube, window_hwidth=1.5)
    first = mmg.pys_moments(cube._data, v)[1]
    b0, b1, b2 = mmg.window_moments(cube, window_hwidth=1.5, v_atpeak=first)
    np.testing.assert_array_equal(a1, b1)


def test_vmask_moments_match_the_spectral_cube_definition():
    cube, cen = _cube(1)
    m0, m1, m2 = mmg.vmask_moments(cube, cen, window_hwidth=1.0)
    v = cube.spectral_axis
    dv = v[1] - v[0]
    for y in range(cube.shape[1]):
        for x in range(cube.shape[2]):
            d = cube._data[:, y, x].astype(np.float64)
            inc = np.isfinite(d) & (v > cen[y, x] - 1.0) & (v < cen[y, x] + 1.0)
            if not inc.any():
                assert np.isnan(m0[y, x])
                continue
            r0, r1, r2 = OG.spectral_cube_moments(v, np.where(inc, d, np.nan))
            np.testing.assert_allclose(m0[y, x], r0, rtol=1e-12)
            np.testing.assert_allclose(m1[y, x], r1, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(m2[y, x], r2, rtol=1e-9)
    assert np.array_equal(mmg.vmask_cube(cube, cen, 1.0)[:, 2, 3], (v > cen[2, 3] - 1.0) & (v < cen[2, 3] + 1.0))


def test_moment_guesses_1c_golden():
    """Against the reference's own moment_guesses_1c (executed by tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "guess_golden.npz"))
    out = mmg.moment_guesses_1c(g["m1c_m0"].copy(), g["m1c_m1"].copy(), g["m1c_m2"].copy())
    np.testing.assert_array_equal(out, g["m1c_out"])
    s = mmg.moment_guesses_1c(float(g["m1c_m0"][1, 2]), float(g["m1c_m1"][1, 2]), float(g["m1c_m2"][1, 2]))
    np.testing.assert_allclose(np.asarray(s, dtype=np.float64), g["m1c_out"][:, 1, 2], rtol=1e-14)


def test_mom_guess_wide_sep_finds_both_lines():
    nchan, ny, nx = 300, 4, 5
    v = (np.arange(nchan) - nchan // 2) * 0.08
    rng = np.random.default_rng(5)
    line = lambda c, a: a * np.exp(-0.5 * ((v[:, None, None] - c) / 0.25) ** 2)   # noqa: E731
    data = line(-0.5, 3.0) + line(3.0, 0.5) + 0.03 * rng.normal(size=(nchan, ny, nx))
    data[:, 0, :] = line(-0.5, 3.0)[:, 0, :] + 0.01 * rng.normal(size=(nchan, nx))   # single line in the first row
    cube = SpecCube(data.astype(np.float32), v)
    mask = np.ones((ny, nx), bool)
    mask[3, 4] = False
    gg = mmg.mom_guess_wide_sep(cube, vpeak=np.full((ny, nx), -0.4), rms=np.full((ny, nx), 0.05), planemask=mask)
    assert gg.shape == (8, ny, nx)
    two = np.zeros((ny, nx), bool)
    two[1:] = True
    two[3, 4] = False
    # component 1 = the flux-weighted centroid of everything within 4 km/s (pulled towards the faint line), component 2
    # = the faint line, found once the emission within two widths of that centroid is blanked
    assert np.all(gg[0][two] > -0.5) and np.all(gg[0][two] < 0.5) and np.abs(gg[4][two] - 3.0).max() < 0.5
    # a single line: the two guesses straddle it, one width apart, each with half the optical depth
    one = np.zeros((ny, nx), bool)
    one[0] = True
    assert np.all(gg[0][one] > gg[4][one]) and np.abs(0.5 * (gg[0] + gg[4])[one] + 0.5).max() < 0.1
    np.testing.assert_allclose(gg[3][one], gg[7][one])
    np.testing.assert_allclose(gg[1][one], gg[5][one])


def test_mad_std_cube_equals_the_numpy_twin():
    from mufasa_b200 import signals
    rng = np.random.default_rng(0)
    for n in (50, 51):                                         # even and odd counts: numpy's median of the middle pair
        d = rng.normal(size=(n, 4, 5)).astype(np.float32)
        d[rng.random(d.shape) < 0.1] = np.nan
        d[:, 0, 0] = np.nan
        a, b = mmg.mad_std_cube(d), signals.mad_std(d, axis=0)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_array_equal(a[np.isfinite(a)], b[np.isfinite(b)])


def test_copy_rows_by_runs():
    from mufasa_b200.dist import copy_rows, owned_rows
    src = np.arange(3 * 40 * 7, dtype=np.float64).reshape(3, 40, 7)
    for rows in (owned_rows(40, 3, 1, "cyclic"), np.array([0, 1, 2, 9, 11, 12, 39]), np.arange(40), np.array([], int)):
        dst = np.full((3, len(rows), 7), -1.0)
        copy_rows(dst, src, rows)
        np.testing.assert_array_equal(dst, src[:, rows])
