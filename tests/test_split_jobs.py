"""Host logic of Engine.fit_batch's job splitting (no GPU): large fits run as independent parts of their pixels, and
parts of a cube that is still being uploaded slab by slab wait only for the slab holding their last pixel."""
import numpy as np
import torch

from mufasa_b200 import _abi
from mufasa_b200.engine import Engine


class _Ev(object):
    def __init__(self, done=False):
        self.done = done

    def query(self):
        return self.done


def _engine():
    eng = Engine.__new__(Engine)            # no CUDA context: only the splitting helpers are exercised
    eng._profiling = False
    return eng


def _job(npix, ncomp, masked=None, slab_ready=None):
    prob = _abi.Problem()
    prob.npix, prob.ncomp = npix, ncomp
    P = 4 * ncomp
    job = dict(prob=prob, spectra=torch.zeros((npix if masked is None else masked[0], 32)), guesses=torch.zeros((npix, P)),
               slab_ready=slab_ready)
    if masked is not None:
        job["pix"] = masked[1]
        job["row_index"] = torch.from_numpy(masked[1])
    out = dict(params=torch.zeros((npix, P)), perr=torch.zeros((npix, P)), chi2=torch.zeros(npix),
               err_used=torch.zeros(npix), status=torch.zeros(npix, dtype=torch.int32), counts=None)
    return job, out


def _cover(parts, npix):
    sizes = [int(p["prob"].npix) for p in parts]
    assert sum(sizes) == npix and all(s > 0 for s in sizes)
    for p in parts:
        n = int(p["prob"].npix)
        assert p["guesses"].shape[0] == n and p["out"]["params"].shape[0] == n and p["out"]["status"].shape[0] == n
    return sizes


def test_generic_split_uses_free_slots_for_the_heaviest_fit():
    eng = _engine()
    (j1, o1), (j2, o2) = _job(65536, 1), _job(65536, 2)
    parts = eng._split_jobs([j1, j2], [o1, o2])
    assert [int(p["prob"].ncomp) for p in parts] == [1, 2, 2]                 # limit 3: the 2c fit in two halves
    assert _cover(parts[1:], 65536) == [32768, 32768]
    # views of the caller's buffers: writing a part's output lands in the full result
    parts[2]["out"]["params"][0, 0] = 7.0
    assert float(o2["params"][32768, 0]) == 7.0
    # a lone fit gets three parts; small fits are left alone; profiled fits run undivided
    assert _cover(eng._split_jobs([j2], [o2]), 65536) == [21845, 21845, 21846]
    js, os_ = _job(20000, 2)
    assert len(eng._split_jobs([js], [os_])) == 1
    eng._profiling = True
    assert len(eng._split_jobs([j1, j2], [o1, o2])) == 2


def test_slab_split_waits_for_the_right_slab():
    eng = _engine()
    ny, nx = 256, 256
    bounds, events = [128 * nx, 256 * nx], [_Ev(), _Ev()]
    (j1, o1), (j2, o2) = _job(ny * nx, 1, slab_ready=(bounds, events)), _job(ny * nx, 2, slab_ready=(bounds, events))
    parts = eng._split_jobs([j1, j2], [o1, o2])
    assert [int(p["prob"].ncomp) for p in parts] == [1, 1, 2, 2]
    assert [p["wait"] for p in parts] == [events[0], events[1], events[0], events[1]]
    assert _cover(parts[:2], ny * nx) == [128 * nx, 128 * nx]
    # masked fit: pixels are cut where they cross the slab boundary; a fit that lives in one slab waits for that slab only
    pix = np.flatnonzero(np.arange(ny * nx) % 4 != 0)[500:]            # 48 652 pixels on both sides of the boundary
    jm, om = _job(pix.size, 2, masked=(ny * nx, pix), slab_ready=(bounds, events))
    parts = eng._split_jobs([jm], [om])
    n0 = int(np.searchsorted(pix, bounds[0]))
    assert _cover(parts, pix.size) == [n0, pix.size - n0]
    assert parts[0]["wait"] is events[0] and parts[1]["wait"] is events[1]
    assert int(parts[0]["row_index"][-1]) < bounds[0] <= int(parts[1]["row_index"][0])
    first_rows = np.flatnonzero(np.arange(ny * nx) < 30000)
    jf, of = _job(first_rows.size, 2, masked=(ny * nx, first_rows), slab_ready=(bounds, events))
    parts = eng._split_jobs([jf], [of])
    assert len(parts) == 1 and parts[0]["wait"] is events[0]
    # five fits do not fit in the batch (8 entries) cut in two: each waits for the slab of its last pixel
    more = [_job(ny * nx, 2, slab_ready=(bounds, events)) for _ in range(3)]
    parts = eng._split_jobs([j1, j2] + [j for j, _ in more], [o1, o2] + [o for _, o in more])
    assert len(parts) == 5 and all(p["wait"] is events[1] for p in parts)
    # once the upload has completed the generic split applies
    for e in events:
        e.done = True
    parts = eng._split_jobs([j1, j2], [o1, o2])
    assert [int(p["prob"].ncomp) for p in parts] == [1, 2, 2] and all(p.get("wait") is None for p in parts)


def test_upload_slab_edges():
    """PCube.slab_edges: the slabs tile the rows in order, the first one is short but holds at least the minimum
    number of spectra a fit part needs, for every cube shape that is split at all (>= 2 slabs)."""
    from mufasa_b200.pcube import PCube
    split_min = 16384
    for ny, nx in ((1024, 1024), (128, 1024), (256, 256), (32, 1024), (53, 700), (2048, 16), (129, 257)):
        nslab = int(max(1, min(4, ny, (ny * nx) // split_min)))
        if nslab < 2:
            continue
        e = PCube.slab_edges(ny, nx, nslab, split_min)
        assert e[0] == 0 and e[-1] == ny and len(e) == nslab + 1
        assert all(b > a for a, b in zip(e[:-1], e[1:]))
        assert e[1] * nx >= split_min and e[1] <= max(ny // 2, -(-split_min // nx))
