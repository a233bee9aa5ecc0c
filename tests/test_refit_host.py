"""Host-side refit-loop helpers (mufasa_b200/master_fitter.py) on small arrays, no GPU."""
import numpy as np

from mufasa_b200 import master_fitter as mf


def test_square_neighbour_and_best_neighbour_lookup():
    s = mf.square_neighbour(1)
    assert s.shape == (3, 3) and s.sum() == 8 and not s[1, 1]
    ref = np.arange(30.0).reshape(5, 6)
    mask = np.zeros((5, 6), bool)
    mask[2, 2] = mask[4, 5] = True
    coords = mf.maxref_neighbor_coords(mask, ref)
    assert tuple(int(c) for c in coords[0]) == (3, 3)           # the largest of the 8 neighbours of (2, 2)
    assert tuple(int(c) for c in coords[1]) == (4, 4)           # neighbours beyond the far edge are skipped
    # the reference only catches IndexError: negative indices wrap around (numpy semantics) -- kept
    mask0 = np.zeros((5, 6), bool)
    mask0[0, 0] = True
    (yy, xx), = mf.maxref_neighbor_coords(mask0, ref)
    assert ref[yy, xx] == ref[-1, -1]


def _neighbour_loop(mask, ref, fill_coord=(0, 0), structure=None):
    """The reference's per-pixel loop (utils/neighbours.py:17-60) restated literally: python max() over the
    neighbours in structure order, IndexError-style skipping of indices past the far edges only."""
    if structure is None:
        structure = mf.square_neighbour(1)
    cy, cx = int(np.ceil(structure.shape[0] / 2)) - 1, int(np.ceil(structure.shape[1] / 2)) - 1
    dy, dx = np.where(structure)
    dy, dx = dy - cy, dx - cx
    ny, nx = ref.shape
    out = []
    for y, x in zip(*np.where(mask)):
        best, best_val = None, None
        for yy, xx in zip(dy + y, dx + x):
            if yy >= ny or xx >= nx or yy < -ny or xx < -nx:
                continue
            val = ref[yy, xx]
            if best is None or val > best_val:
                best, best_val = (yy, xx), val
        out.append(best if best is not None else fill_coord)
    return out


def test_best_neighbour_lookup_matches_the_sequential_loop():
    """NaN and -inf reference values, ties, wrap-around at the near edges, empty structures, 1-pixel maps."""
    rng = np.random.default_rng(1)
    for trial in range(25):
        ny, nx = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        ref = rng.integers(0, 4, size=(ny, nx)).astype(float)
        ref[rng.random((ny, nx)) < 0.3] = np.nan
        if trial % 3 == 0:
            ref[rng.random((ny, nx)) < 0.2] = -np.inf
        mask = rng.random((ny, nx)) < 0.6
        for st in (None, mf.square_neighbour(2), mf.disk_neighbour(2), np.zeros((3, 3), bool)):
            want = _neighbour_loop(mask, ref, structure=st)
            got = mf.maxref_neighbor_coords(mask, ref, structure=st)
            assert [tuple(int(v) for v in c) for c in got] == [tuple(int(v) for v in c) for c in want]


class _P(object):
    def __init__(self, parcube):
        self.parcube = parcube


class _U(object):
    def __init__(self, parcube):
        self.pcubes = {"1": _P(parcube)}


def test_get_refit_guesses_best_neighbour_and_convolved():
    ny, nx = 6, 7
    par = np.zeros((4, ny, nx))
    par[:, 1:5, 1:6] = np.arange(1, 5)[:, None, None] + 0.01 * np.arange(nx - 2)[None, None, :]
    mask = np.zeros((ny, nx), bool)
    mask[2, 3] = True
    mask[0, 0] = True                                          # no fitted neighbour -> dropped from the mask
    ref = np.zeros((ny, nx))
    ref[3, 4] = 9.0
    g, m = mf.get_refit_guesses(_U(par.copy()), mask, ncomp=1, refmap=ref)
    assert np.array_equal(g[:, 2, 3], par[:, 3, 4])
    assert m[2, 3] and not m[0, 0]
    g2, m2 = mf.get_refit_guesses(_U(par.copy()), mask, ncomp=1, method='convolved')
    assert np.isfinite(g2[:, 2, 3]).all() and abs(g2[0, 2, 3] - par[0, 2, 2:5].mean()) < 0.05


def test_marginal_pixels_and_local_bad():
    lnk = np.full((20, 20), np.nan)
    lnk[5:15, 5:15] = 20.0
    lnk[9, 9] = 1.0                                            # a hole inside the structure
    lnk[0, 0] = 30.0                                           # an isolated pixel: too small to be a structure
    rim = mf.get_marginal_pix(lnk, lnk_thresh=5)
    assert rim.sum() == 41 and rim[9, 9] and rim[4, 7] and not rim[4, 4] and not rim[0, 1]
    holes = mf.get_marginal_pix(lnk, lnk_thresh=5, holes_only=True)
    assert holes.sum() == 1 and holes[9, 9]
    d = mf.disk_neighbour(2)
    assert d.shape == (5, 5) and d.sum() == 12 and not d[2, 2]
    # get_local_bad: a small dip inside a bright plateau is flagged, the plateau is not
    yy, xx = np.mgrid[:40, :40]
    lnk = 30.0 + 0.0 * yy
    lnk[18:21, 18:21] = 12.0
    bad = mf.get_local_bad(lnk, lnk_thresh=10, block_size=15, bad_size_max=225)
    assert bad[19, 19] and bad.sum() == 9


def test_masked_moment_of_the_data_over_the_model_footprint():
    from mufasa_b200.UltraCube import get_masked_moment
    from mufasa_b200.cube import SpecCube
    nchan, ny, nx = 60, 4, 5
    v = np.linspace(-3, 3, nchan)
    rng = np.random.default_rng(0)
    data = rng.normal(0, 0.1, (nchan, ny, nx)).astype(np.float32)
    model = np.zeros((nchan, ny, nx))
    model[28:32] = 2.0                                         # a bright line everywhere ...
    model[:, 0, 0] = 0.0
    model[29:31, 0, 0] = 0.05                                  # ... but one faint pixel
    data = data + model.astype(np.float32)
    cube = SpecCube(data, v)
    m0 = get_masked_moment(cube, model.copy(), order=0, expand=4)
    dv = v[1] - v[0]
    chan = np.zeros(nchan, bool)
    chan[28 - 2:32 + 2] = True                                 # ones(4) dilation: [k - 2, k + 1]
    chan[32 + 1] = False
    want = data[chan, 1, 1].astype(np.float64).sum() * dv
    assert np.isclose(m0[1, 1], want)
    # the faint pixel borrows the channels where any model exceeds 10 % of the median peak
    assert np.isclose(m0[0, 0], data[chan, 0, 0].astype(np.float64).sum() * dv)
