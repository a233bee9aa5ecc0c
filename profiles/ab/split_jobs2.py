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
js = []
for nc in (1, 2):
    lo = [-3.5, mm.sigmin, mm.Texmin, mm.taumin] * nc
    hi = [4.5, mm.sigmax, mm.Texmax, mm.taumax] * nc
    gg = synth.clamp_guesses(syn["guess%d" % nc], lo, hi, eps=mm.eps)
    g = torch.from_numpy(np.ascontiguousarray(gg.reshape(4 * nc, npix).T)).to(dev)
    js.append(dict(prob=_abi.make_problem(fittype, nc, nchan, npix, v[0], dv, lo, hi), spectra=rows, guesses=g))
ref = None
for limit in (2, 3, 4, 2, 3, 4):
    os.environ["B200FIT_SPLIT_LIMIT"] = str(limit)
    ts = []
    for r in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs = eng.fit_batch(js)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    sig = [float(o["params"].sum()) for o in outs] + [int(o["status"].sum()) for o in outs]
    if ref is None: ref = sig
    print("jobs <= %d: %s ms  identical=%s" % (limit, ["%.2f" % t for t in ts], sig == ref), flush=True)
for limit in (1, 2, 3, 4):
    os.environ["B200FIT_SPLIT_LIMIT"] = str(limit)
    ts = []
    for r in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs = eng.fit_batch(js[1:])
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("2c alone, jobs <= %d: %s ms" % (limit, ["%.2f" % t for t in ts]), flush=True)
