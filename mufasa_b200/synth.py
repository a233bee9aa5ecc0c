"""Synthetic NH3 / N2H+ cubes of BASELINE.json's configs.

The reference ships no test data, so the workloads are defined here.  The generator does not evaluate the
spectral model itself: the caller passes ``modelcube_fn(parcube[P, ny, nx], v_kms, fittype) -> [nchan, ny, nx]``
(the GPU's fp64 ``b200fit_model`` in bench.py; the CPU oracle in the CPU tests).  Noise and guess
perturbations are drawn per row from independent child streams of ``SeedSequence(20261019 + cfg)`` so any
row slab (multi-GPU partition) is reproducible on its own.
"""
import numpy as np

SEED0 = 20261019
NOISE_RMS = 0.1

CONFIGS = {
    # cfg: (fittype, ny, nx, nchan, dv)
    1: ("nh3_multi_v", 64, 64, 500, 0.10),
    2: ("nh3_multi_v", 256, 256, 800, 0.0724),
    3: ("nh3_multi_v", 512, 512, 1000, 0.0724),
    4: ("n2hp_multi_v", 512, 512, 1200, 0.05),
    5: ("nh3_multi_v", 1024, 1024, 1000, 0.0724),
}


def vaxis(nchan, dv):
    """V_i = (i - nchan/2) * dv  (km/s, ascending)."""
    return (np.arange(nchan, dtype=np.float64) - nchan // 2) * dv


def truth_fields(ny, nx, rows=None):
    """Smooth truth parameter fields on x^, y^ in [-1, 1].  Returns (par1[4,ry,nx], par2[4,ry,nx])."""
    y0, y1 = (0, ny) if rows is None else rows
    yh = (np.arange(y0, y1, dtype=np.float64) / max(ny - 1, 1) * 2.0 - 1.0)[:, None]
    xh = (np.arange(nx, dtype=np.float64) / max(nx - 1, 1) * 2.0 - 1.0)[None, :]
    yh = np.broadcast_to(yh, (y1 - y0, nx))
    xh = np.broadcast_to(xh, (y1 - y0, nx))
    v1 = 1.5 * xh + 0.4 * np.sin(3 * np.pi * yh)
    s1 = 0.12 + 0.24 * (1 + xh * yh)
    t1 = 6.5 + 2.5 * yh
    a1 = 3.25 + 2.75 * np.sin(2 * np.pi * xh)
    v2 = v1 + (1.0 + 0.6 * np.cos(2 * np.pi * yh))
    par1 = np.stack([v1, s1, t1, a1])
    par2 = np.stack([v2, 0.8 * s1, 0.8 * t1, 0.4 * a1])
    return par1, par2


def ncomp_truth(cfg, ny, nx, rows=None):
    """Known number of components per pixel."""
    y0, y1 = (0, ny) if rows is None else rows
    yy, xx = np.meshgrid(np.arange(y0, y1), np.arange(nx), indexing="ij")
    if cfg == 1:
        return np.ones((y1 - y0, nx), dtype=np.int8)
    if cfg == 4:
        return np.full((y1 - y0, nx), 2, dtype=np.int8)
    nc = np.ones((y1 - y0, nx), dtype=np.int8)
    r = np.hypot(yy - (ny - 1) / 2.0, xx - (nx - 1) / 2.0)
    nc[r < 0.35 * ny] = 2
    border = 16 if min(ny, nx) >= 64 else 0
    if border:
        edge = (yy < border) | (yy >= ny - border) | (xx < border) | (xx >= nx - border)
        nc[edge] = 0
    return nc


def amplitude(cfg, ny, nx, rows=None):
    """cfg 4: amplitude ramp so that peak SNR spans ~1..20 with about a third of the pixels below 3."""
    y0, y1 = (0, ny) if rows is None else rows
    if cfg != 4:
        return np.ones((y1 - y0, nx))
    yy, xx = np.meshgrid(np.arange(y0, y1), np.arange(nx), indexing="ij")
    t = (yy + xx) / float(ny + nx - 2)
    return 0.05 + 1.0 * t ** 2.2


def make_cube(cfg, modelcube_fn, rows=None, shape=None, dtype=np.float32, seed_offset=0, out=None):
    """Build (a row slab of) the synthetic cube of config ``cfg``.

    Returns dict(cube[nchan,ry,nx] float32, v[nchan], fittype, dv, truth1, truth2, ncomp_true,
                 guess1[4,ry,nx], guess2[8,ry,nx]).  ``shape=(ny,nx,nchan)`` overrides the config size
    (reduced-size parity cases keep the config's species / channel width).  ``seed_offset`` selects an independent
    noise realisation (one per rank in weak-scaling runs); ``out`` = preallocated [nchan, ry, nx] buffer for the
    cube (e.g. a view of pinned memory)."""
    fittype, ny, nx, nchan, dv = CONFIGS[cfg]
    if shape is not None:
        ny, nx, nchan = shape
    y0, y1 = (0, ny) if rows is None else rows
    v = vaxis(nchan, dv)
    par1, par2 = truth_fields(ny, nx, (y0, y1))
    nct = ncomp_truth(cfg, ny, nx, (y0, y1))
    amp = amplitude(cfg, ny, nx, (y0, y1))
    m2 = modelcube_fn(np.concatenate([par1, par2]), v, fittype)
    m1 = modelcube_fn(par1, v, fittype)
    model = np.where(nct[None] == 2, m2, np.where(nct[None] == 1, m1, 0.0)) * amp[None]
    children = np.random.SeedSequence(SEED0 + cfg + 1000 * int(seed_offset)).spawn(2 * ny)
    cube = out if out is not None else np.empty((nchan, y1 - y0, nx), dtype=dtype)
    g1 = np.empty((4, y1 - y0, nx))
    g2 = np.empty((8, y1 - y0, nx))
    truth = np.concatenate([par1, par2])
    for r, y in enumerate(range(y0, y1)):
        rng = np.random.Generator(np.random.PCG64(children[y]))
        cube[:, r, :] = (model[:, r, :] + rng.normal(0.0, NOISE_RMS, size=(nchan, nx))).astype(dtype)
        rg = np.random.Generator(np.random.PCG64(children[ny + y]))
        pert = 1.0 + 0.15 * rg.normal(size=(8, nx))
        dvel = 0.1 * rg.normal(size=(2, nx))
        g = truth[:, r, :] * pert
        g[0] = truth[0, r, :] + dvel[0]
        g[4] = truth[4, r, :] + dvel[1]
        if cfg == 4:   # 10 % of the pixels start with tau exactly on its lower limit (limit stress)
            pick = rg.random(nx) < 0.10
            g[3, pick] = 0.1
        g2[:, r, :] = g
        g1[:, r, :] = g[:4]
    return dict(cube=cube, v=v, fittype=fittype, dv=dv, truth1=par1, truth2=truth, ncomp_true=nct,
                guess1=g1, guess2=g2, amp=amp)


def clamp_guesses(guesses, lo, hi, eps=1e-3):
    """Keep perturbed-truth guesses inside the box the way cubefit_simp does (multi_v_fit.py:178-190)."""
    g = guesses.copy()
    for j in range(g.shape[0]):
        lo_j, hi_j = lo[j], hi[j]
        g[j][g[j] < lo_j] = lo_j + eps
        g[j][g[j] > hi_j] = hi_j - eps
    return g
