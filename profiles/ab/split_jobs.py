"""Experiment: the 2c fit of 65 536 spectra as 1, 2 or 3 jobs of a batch (with the 1c fit alongside)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from mufasa_b200 import _abi, synth
from mufasa_b200.engine import get_engine
from mufasa_b200.spec_models.meta_model import MetaModel
eng = get_engine(0); dev = eng.device
fittype, ny, nx, nchan, dv = synth.CONFIGS[2]
def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].cpu().numpy().T.reshape(len(v), ry, rx)
syn = synth.make_cube(2, gpu_modelcube)
v = syn["v"]; npix = ny * nx
mm = MetaModel(fittype, 1)
rows = eng.gather(torch.from_numpy(syn["cube"].reshape(nchan, npix)).to(dev))
def problem(nc, n):
    lo = [-3.5, mm.sigmin, mm.Texmin, mm.taumin] * nc
    hi = [4.5, mm.sigmax, mm.Texmax, mm.taumax] * nc
    return _abi.make_problem(fittype, nc, nchan, n, v[0], dv, lo, hi), lo, hi
g = {}
for nc in (1, 2):
    _, lo, hi = problem(nc, npix)
    gg = synth.clamp_guesses(syn["guess%d" % nc], lo, hi, eps=mm.eps)
    g[nc] = torch.from_numpy(np.ascontiguousarray(gg.reshape(4 * nc, npix).T)).to(dev)
def jobs(nsplit, interleave):
    out = [dict(prob=problem(1, npix)[0], spectra=rows, guesses=g[1])]
    if nsplit == 1:
        out.append(dict(prob=problem(2, npix)[0], spectra=rows, guesses=g[2]))
        return out
    idx = torch.arange(npix, device=dev)
    for k in range(nsplit):
        sel = idx[k::nsplit] if interleave else idx[k * npix // nsplit:(k + 1) * npix // nsplit]
        sel = sel.contiguous()
        out.append(dict(prob=problem(2, sel.numel())[0], spectra=rows, guesses=g[2][sel].contiguous(), row_index=sel))
    return out
for nsplit, inter in ((1, False), (2, False), (2, True), (3, True), (1, False)):
    js = jobs(nsplit, inter)
    ts = []
    for r in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs = eng.fit_batch(js)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("2c as %d job(s)%s + 1c: %s ms" % (nsplit, " interleaved" if inter else "", ["%.2f" % t for t in ts]), flush=True)
