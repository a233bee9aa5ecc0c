#!/usr/bin/env python
"""BASELINE configs[2]-style pipeline on one GPU: UltraCube.fit_cube([1, 2]) through cubefit_gen (GPU moment pre-pass
-> moment_guesses -> tautex_renorm -> clamps -> masks -> batched 1c + 2c fits) + get_best_2c_parcube, on a row slab
of the 512 x 512 x 1000 NH3 cube.   usage: python profiles/pipeline_cfg3.py [nrows]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mufasa_b200 import _abi, synth                      # noqa: E402
from mufasa_b200.cube import SpecCube                     # noqa: E402
from mufasa_b200.engine import get_engine                 # noqa: E402
from mufasa_b200.UltraCube import UltraCube               # noqa: E402

nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = get_engine(0)
dev = eng.device
fittype, ny, nx, nchan, dv = synth.CONFIGS[3]


def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].cpu().numpy().T.reshape(len(v), ry, rx)


y0 = (ny - nrows) // 2
pinned = torch.empty((nchan, nrows, nx), dtype=torch.float32).pin_memory()
syn = synth.make_cube(3, gpu_modelcube, rows=(y0, y0 + nrows), out=pinned.numpy())
spec = SpecCube(syn["cube"], syn["v"])


def step(snr_min=3.0):
    uc = UltraCube(cube=spec, fittype=fittype, device=0)
    uc.fit_cube(ncomp=[1, 2], snr_min=snr_min)
    res = uc.get_best_2c_parcube()
    torch.cuda.synchronize()
    return uc


step()
t0 = time.perf_counter()
uc = step()
dt = time.perf_counter() - t0
nfit = int(uc.pcubes["2"].has_fit.sum())
print("pipeline on %d x %d x %d: %.1f ms end to end (host buffers in, maps out); %d pixels fitted (snr_min 3); "
      "%.0f fitted spectra/s" % (nrows, nx, nchan, dt * 1e3, nfit, nfit / dt))
truth = syn["ncomp_true"]
print("ncomp map vs truth: %.3f agree; hist %s" % ((uc.ncomp_map == truth).mean(),
                                                  np.bincount(uc.ncomp_map.ravel().astype(int), minlength=3).tolist()))
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
