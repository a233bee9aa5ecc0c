"""GPU parity of the moment / signal-mask pre-pass: bit-cube morphology vs scipy.ndimage with the
structuring elements and border conventions of oracle/signals.py, and the composed get_moments vs the numpy
restatement of mufasa/signals.py:329-396 in oracle/signals.py."""
import numpy as np
import pytest

import _tiles as T
from mufasa_b200 import prepass, synth
from oracle import signals
from mufasa_b200.cube import SpecCube

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    from mufasa_b200.engine import get_engine
    assert torch.cuda.is_available()
    return get_engine(0)


def _to_bits(mask):
    """bool [nchan, ny, nx] -> int32 [ny*nx, W] (bit b of word w = channel 32 w + b)."""
    nchan, ny, nx = mask.shape
    W = (nchan + 31) // 32
    padded = np.zeros((W * 32, ny * nx), dtype=np.uint64)
    padded[:nchan] = mask.reshape(nchan, -1)
    words = (padded.reshape(W, 32, ny * nx) << np.arange(32, dtype=np.uint64)[None, :, None]).sum(axis=1)
    return np.ascontiguousarray(words.T.astype(np.uint32)).view(np.int32)


def _from_bits(bits, nchan, ny, nx):
    b = bits.view(np.uint32)
    W = b.shape[1]
    chan = ((b[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(ny * nx, W * 32)[:, :nchan]
    return chan.T.reshape(nchan, ny, nx).astype(bool)


@pytest.mark.parametrize("nchan,ny,nx", [(70, 9, 11), (64, 5, 4), (33, 7, 7)])
def test_bit_cube_morphology_matches_scipy(engine, nchan, ny, nx):
    import torch
    rng = np.random.default_rng(nchan)
    mask = rng.random((nchan, ny, nx)) < 0.8          # dense, so that erosions keep something
    mask[:, ny // 2, nx // 2] = True
    W = (nchan + 31) // 32
    rows = torch.zeros((ny * nx, W * 32), dtype=torch.float32, device="cuda")
    ops = prepass.BitCubeOps(engine, rows, nchan, ny, nx, 0.0, 1.0)
    d_bits = torch.from_numpy(_to_bits(mask)).cuda()
    assert np.array_equal(_from_bits(d_bits.cpu().numpy(), nchan, ny, nx), mask)       # the test's own packing
    cases = [(prepass.MORPH_ERODE_BALL2, signals.binary_erosion(mask, signals.ball(2))),
             (prepass.MORPH_ERODE_CROSS, signals.binary_erosion(mask)),
             (prepass.MORPH_DILATE_CROSS, signals.binary_dilation(mask & (rng.random(mask.shape) < 0.05))),
             (prepass.MORPH_DILATE_DISK3, None)]
    sparse = mask & (rng.random(mask.shape) < 0.03)
    for op, ref in cases:
        src = mask
        if op == prepass.MORPH_DILATE_CROSS:
            src = mask & (np.random.default_rng(1).random(mask.shape) < 0.05)
            ref = signals.binary_dilation(src)
        if op == prepass.MORPH_DILATE_DISK3:
            src = sparse
            ref = signals.binary_dilation(src, signals.disk(3)[np.newaxis, :])
        out = ops.morph(op, torch.from_numpy(_to_bits(src)).cuda()).cpu().numpy()
        assert np.array_equal(_from_bits(out, nchan, ny, nx), ref), op
    opened = ops.opening(d_bits).cpu().numpy()
    assert np.array_equal(_from_bits(opened, nchan, ny, nx), signals.binary_opening(mask))


@pytest.mark.parametrize("cfg,central", [(1, None), (4, 3.5)])
def test_get_moments_matches_numpy_mirror(cfg, central):
    from mufasa_b200.pcube import PCube
    shape = (20, 22, 500) if cfg == 1 else (18, 20, 1200)
    syn = synth.make_cube(cfg, T.oracle_modelcube, shape=shape)
    cube = syn["cube"].copy()
    cube[:, 0, :] = np.nan                      # an empty border row
    cube[40:60, 5, 5] = np.nan
    sc = SpecCube(cube, syn["v"])
    include = signals.trim_cube_edge(sc, trim=3, min_size=100)
    planemask = np.any(include, axis=0)
    with np.errstate(all="ignore"):
        ref = signals.get_moments(sc, window_hwidth=5 if cfg == 1 else 4, linewidth_sigma=False, trim=None,
                                  return_rms=True, central_win_hwidth=central, include=include)
    p = PCube(cube, syn["v"], device=0)
    got = prepass.get_moments(p, planemask, window_hwidth=5 if cfg == 1 else 4, linewidth_sigma=False,
                              central_win_hwidth=central, return_rms=True)
    names = ("m0", "m1", "m2", "rms")
    for name, a, b in zip(names, got, ref):
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = np.isfinite(b)
        assert ok.sum() > 20, name              # config 4 is the low-S/N cube: most pixels have no 3-sigma signal
        if name == "rms":
            assert np.array_equal(a[ok], b[ok])
        else:
            # same channels enter the sums (the masks are integer work); fp64 sums in a different order
            assert np.allclose(a[ok], b[ok], rtol=1e-9, atol=1e-12), (name, np.max(np.abs(a[ok] - b[ok])))
