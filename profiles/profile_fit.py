#!/usr/bin/env python
"""Profiling workload: a row slab of BASELINE configs[1] (NH3 256x256x800) through the C ABI on one GPU:
1-comp fit, 2-comp fit, AICc selection.  Used under ncu (launch list + --set full of fit_kernel<2>).
usage: python profiles/profile_fit.py [nrows] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mufasa_b200 import _abi, synth                      # noqa: E402
from mufasa_b200.engine import get_engine                 # noqa: E402
from mufasa_b200.spec_models.meta_model import MetaModel  # noqa: E402

nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = int(os.environ.get("PROFILE_CFG", "2"))
eng = get_engine(0)
dev = eng.device
fittype, ny, nx, nchan, dv = synth.CONFIGS[cfg]


def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
    m = eng.model(prob, par)[:, :len(v)]
    return m.cpu().numpy().T.reshape(len(v), ry, rx)


y0 = (ny - nrows) // 2
syn = synth.make_cube(cfg, gpu_modelcube, rows=(y0, y0 + nrows))
v = syn["v"]
npix = nrows * nx
mm = MetaModel(fittype, 1)
cube_dev = torch.from_numpy(syn["cube"].reshape(nchan, npix)).to(dev)
rows = eng.gather(cube_dev)
probs, gs = {}, {}
for nc in (1, 2):
    vlo, vhi = (-3.5, 4.5) if os.environ.get("PROFILE_VLIM", "narrow") == "narrow" else (-11.0, 11.0)
    lo = [vlo, mm.sigmin, mm.Texmin, mm.taumin] * nc
    hi = [vhi, mm.sigmax, mm.Texmax, mm.taumax] * nc
    g = synth.clamp_guesses(syn["guess%d" % nc], lo, hi, eps=mm.eps)
    gs[nc] = torch.from_numpy(np.ascontiguousarray(g.reshape(4 * nc, npix).T)).to(dev)
    probs[nc] = _abi.make_problem(fittype, nc, nchan, npix, v[0], dv, lo, hi)
if os.environ.get("PROFILE_EVENTS", "0") == "1":
    eng.set_profiling(True)
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o1 = eng.fit(probs[1], rows, gs[1])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if os.environ.get("PROFILE_EVENTS", "0") == "1":
        print("   1c profile:", eng.get_profile(), flush=True)
    o2 = eng.fit(probs[2], rows, gs[2])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if os.environ.get("PROFILE_EVENTS", "0") == "1":
        print("   2c profile:", eng.get_profile(), flush=True)
    a = eng.aicc_global(probs[1], rows, o1["params"], o2["params"])
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    st = o2["status"].cpu().numpy()
    cn = o2["counts"].cpu().numpy().astype(np.float64)
    print("rep %d npix %d: fit1 %.2f ms  fit2 %.2f ms  aicc %.2f ms | 2c: conv %.3f maxiter %.3f mean evals %.1f "
          "(median %.0f, p90 %.0f, max %.0f) | ncomp hist %s" % (
              r, npix, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, np.isin(st, [1, 2, 3, 4, 6]).mean(),
              (st == 5).mean(), cn[:, 0].mean(), np.median(cn[:, 0]), np.percentile(cn[:, 0], 90), cn[:, 0].max(),
              np.bincount(a["ncomp"].cpu().numpy().astype(int), minlength=3).tolist()), flush=True)
    if os.environ.get("PROFILE_BATCH", "0") == "1":
        torch.cuda.synchronize()
        tb0 = time.perf_counter()
        ob = eng.fit_batch([dict(prob=probs[1], spectra=rows, guesses=gs[1]),
                            dict(prob=probs[2], spectra=rows, guesses=gs[2])])
        torch.cuda.synchronize()
        tb1 = time.perf_counter()
        same = all(bool(torch.equal(ob[0][k], o1[k])) and bool(torch.equal(ob[1][k], o2[k]))
                   for k in ("params", "perr", "chi2", "status"))
        print("   batch(1c || 2c): %.2f ms   identical to the separate fits: %s" % ((tb1 - tb0) * 1e3, same), flush=True)
    if os.environ.get("PROFILE_BATCH", "0") == "2":
        half = npix // 2
        jobs = []
        for nc in (1, 2):
            for (a, b) in ((0, half), (half, npix)):
                pj = _abi.make_problem(fittype, nc, nchan, b - a, v[0], dv, [float(x) for x in probs[nc].lo[:4 * nc]],
                                       [float(x) for x in probs[nc].hi[:4 * nc]])
                jobs.append(dict(prob=pj, spectra=rows[a:b], guesses=gs[nc][a:b].contiguous()))
        for _ in range(2):
            torch.cuda.synchronize()
            tb0 = time.perf_counter()
            ob = eng.fit_batch(jobs)
            torch.cuda.synchronize()
            tb1 = time.perf_counter()
        same = bool(torch.equal(torch.cat([ob[2]["params"], ob[3]["params"]]), o2["params"])) and \
            bool(torch.equal(torch.cat([ob[0]["params"], ob[1]["params"]]), o1["params"]))
        print("   batch of 4 (1c, 2c) x (two pixel halves): %.2f ms   identical: %s" % ((tb1 - tb0) * 1e3, same), flush=True)
    if r == 0:
        ne = cn[:, 0]
        print("   active pixels at round r (2c):", {rr: int((ne > rr).sum()) for rr in (0, 5, 10, 20, 40, 80, 120, 160, 200, 240, 280, 320)}, flush=True)
    hist = np.bincount(st.clip(-1, 8) + 1, minlength=10)
    print("status hist (-1..8):", hist.tolist(), flush=True)
