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
