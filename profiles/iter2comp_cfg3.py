#!/usr/bin/env python
"""BASELINE configs[2] (512 x 512 x 1000 NH3 cube) through the first stage of the reference's full pipeline,
Region + iter_2comp_fit (master_fitter.py:65-126, :682-775), from a FITS file on disk: convolve to 2 x the beam and
regrid, fit the convolved cube (1c + 2c, moment guesses), regrid its parameters into guesses, fit the full-resolution
cube (1c, 2c), save every product, then the AICc selection.  Wall-clock per stage.
usage: python profiles/iter2comp_cfg3.py [ny nx nchan] [workdir]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mufasa_b200 import _abi, synth                      # noqa: E402
from mufasa_b200 import convolve_tools as ct             # noqa: E402
from mufasa_b200 import master_fitter as mf              # noqa: E402
from mufasa_b200.cube import SpecCube                     # noqa: E402
from mufasa_b200.engine import get_engine                 # noqa: E402

fittype, ny, nx, nchan, dv = synth.CONFIGS[3]
if len(sys.argv) >= 4:
    ny, nx, nchan = (int(a) for a in sys.argv[1:4])
work = sys.argv[4] if len(sys.argv) > 4 else tempfile.mkdtemp(prefix="iter2comp_")
eng = get_engine(0)
dev = eng.device


def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].cpu().numpy().T.reshape(len(v), ry, rx)


T = {}


def lap(name, t0):
    torch.cuda.synchronize()
    T[name] = time.perf_counter() - t0
    print("%-34s %8.1f ms" % (name, T[name] * 1e3), flush=True)
    return time.perf_counter()


t0 = time.perf_counter()
syn = synth.make_cube(3, gpu_modelcube, shape=(ny, nx, nchan))
hdr = {"CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 52.0, "CRVAL2": 31.0, "CUNIT1": "deg", "CUNIT2": "deg",
       "CDELT1": -0.0024, "CDELT2": 0.0024, "CRPIX1": nx / 2.0, "CRPIX2": ny / 2.0, "NAXIS1": nx, "NAXIS2": ny,
       "BMAJ": 0.0088, "BMIN": 0.0088, "BPA": 0.0}
path = os.path.join(work, "field.fits")
SpecCube(syn["cube"], syn["v"], rest_value=23.6944955e9, header=hdr).write(path)
t0 = lap("synthesise + write FITS (setup)", t0)

reg = mf.Region(path, "field", paraDir=os.path.join(work, "paras"), fittype=fittype, device=0)
t0 = lap("Region: read FITS cube", t0)
reg.get_convolved_cube(update=True)
t0 = lap("convolve x2 + regrid + write + reopen", t0)
reg.get_convolved_fits([1, 2], update=True, snr_min=3.0)
t0 = lap("fit convolved cube (1c, 2c) + save", t0)
if os.environ.get("PROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    mf.iter_2comp_fit(reg, snr_min=3.0, updateCnvFits=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
else:
    mf.iter_2comp_fit(reg, snr_min=3.0, updateCnvFits=False)
t0 = lap("guesses from cnv fits + full-res fits", t0)
best = reg.ucube.get_best_2c_parcube()
t0 = lap("AICc selection", t0)

p1, p2 = reg.ucube.pcubes["1"], reg.ucube.pcubes["2"]
truth = syn["ncomp_true"]
fit = p1.has_fit
one = fit & (truth == 1)
print("cube %d x %d x %d; fitted pixels 1c %d 2c %d of %d" % (ny, nx, nchan, p1.has_fit.sum(), p2.has_fit.sum(), ny * nx))
print("1c vlsr median |fit - truth| on 1-component pixels: %.4f km/s" % np.median(np.abs(p1.parcube[0][one] - syn["truth1"][0][one])))
print("ncomp map vs truth on fitted pixels: %.3f agree; hist %s" % (
    (reg.ucube.ncomp_map[fit] == truth[fit]).mean(), np.bincount(reg.ucube.ncomp_map.ravel().astype(int), minlength=3).tolist()))
total = sum(v for k, v in T.items() if "setup" not in k)
print("pipeline total (without setup): %.1f ms" % (total * 1e3))
for f in sorted(os.listdir(work)) + sorted(os.listdir(os.path.join(work, "paras"))):
    print("  product:", f)
