#!/usr/bin/env python
"""BASELINE configs[2] -- "512 x 512 x 1000 full pipeline" -- through the stages of master_fitter.master_2comp_fit
from a FITS file on disk: two-step fit through the convolved cube, bad-pixel, wide-separation (reference behaviour, and
with WIDE_AS_DOCUMENTED=1 the documented masking) and marginal refits, expansion into the surroundings, products.  Wall-clock per stage.   usage: python profiles/master2comp_cfg3.py [ny nx nchan]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mufasa_b200 import _abi, synth                      # noqa: E402
from mufasa_b200 import master_fitter as mf              # noqa: E402
from mufasa_b200.cube import SpecCube                     # noqa: E402
from mufasa_b200.engine import get_engine                 # noqa: E402

fittype, ny, nx, nchan, dv = synth.CONFIGS[3]
if len(sys.argv) >= 4:
    ny, nx, nchan = (int(a) for a in sys.argv[1:4])
work = tempfile.mkdtemp(prefix="master2comp_")
eng = get_engine(0)
dev = eng.device


def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].cpu().numpy().T.reshape(len(v), ry, rx)


def lap(name, t0):
    torch.cuda.synchronize()
    print("%-44s %9.1f ms" % (name, (time.perf_counter() - t0) * 1e3), flush=True)
    return time.perf_counter()


t0 = time.perf_counter()
syn = synth.make_cube(3, gpu_modelcube, shape=(ny, nx, nchan))
hdr = {"CTYPE1": "RA---TAN", "CTYPE2": "DEC--TAN", "CRVAL1": 52.0, "CRVAL2": 31.0, "CUNIT1": "deg", "CUNIT2": "deg",
       "CDELT1": -0.0024, "CDELT2": 0.0024, "CRPIX1": nx / 2.0, "CRPIX2": ny / 2.0, "NAXIS1": nx, "NAXIS2": ny,
       "BMAJ": 0.0088, "BMIN": 0.0088, "BPA": 0.0}
path = os.path.join(work, "field.fits")
SpecCube(syn["cube"], syn["v"], rest_value=23.6944955e9, header=hdr).write(path)
t0 = lap("synthesise + write FITS (setup)", t0)
tstart = time.perf_counter()
reg = mf.Region(path, "field", paraDir=os.path.join(work, "paras"), fittype=fittype, device=0)
t0 = lap("Region: read FITS cube", t0)
snr_min = 3.0
mf.iter_2comp_fit(reg, snr_min=snr_min)
t0 = lap("iter_2comp_fit (convolve, cnv fits, full fits)", t0)
mf.refit_bad_2comp(reg, snr_min=1.0, lnk_thresh=-5)
t0 = lap("refit_bad_2comp", t0)
mf.RESIDUAL_MASK_AS_DOCUMENTED = bool(os.environ.get("WIDE_AS_DOCUMENTED"))
good = mf.refit_2comp_wide(reg, snr_min=1.0)
t0 = lap("refit_2comp_wide (%s; refits kept: %s)" % (
    "documented masking" if mf.RESIDUAL_MASK_AS_DOCUMENTED else "reference branches", None if good is None else int(good.sum())), t0)
for nc in (1, 2):
    mf.refit_marginal(reg, ncomp=nc, lnk_thresh=10, holes_only=False, method='best_neighbour')
t0 = lap("refit_marginal (1c, 2c)", t0)
if os.environ.get("PROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    mf.fit_surroundings(reg, ncomps=[1, 2], snr_min=snr_min, max_iter=None, match_footprint=True)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
else:
    mf.fit_surroundings(reg, ncomps=[1, 2], snr_min=snr_min, max_iter=None, match_footprint=True)
t0 = lap("fit_surroundings", t0)
mf.refit_bad_2comp(reg, snr_min=1.0, lnk_thresh=-5)
t0 = lap("refit_bad_2comp (second pass)", t0)
mf.save_best_2comp_fit(reg)
t0 = lap("save_best_2comp_fit (products)", t0)
print("pipeline total (without setup): %.1f s" % (time.perf_counter() - tstart))
for e in reg.progress_log:
    if e.get("finished"):
        print("   ", e)
p1 = reg.ucube.pcubes["1"]
truth = syn["ncomp_true"]
fit = p1.has_fit
print("cube %d x %d x %d; fitted pixels %d of %d; ncomp map vs truth on fitted pixels: %.3f; hist %s" % (
    ny, nx, nchan, fit.sum(), ny * nx, (reg.ucube.ncomp_map[fit] == truth[fit]).mean(),
    np.bincount(reg.ucube.ncomp_map.ravel().astype(int), minlength=3).tolist()))
print("products:", sorted(os.listdir(os.path.join(work, "paras"))))
