import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
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
v = syn["v"]; npix = ny * nx; mm = MetaModel(fittype, 1)
rows = eng.gather(torch.from_numpy(syn["cube"].reshape(nchan, npix)).to(dev))
outs = {}
for nc in (1, 2):
    lo = [-3.5, mm.sigmin, mm.Texmin, mm.taumin] * nc; hi = [4.5, mm.sigmax, mm.Texmax, mm.taumax] * nc
    g = synth.clamp_guesses(syn["guess%d" % nc], lo, hi, eps=mm.eps)
    gd = torch.from_numpy(np.ascontiguousarray(g.reshape(4 * nc, npix).T)).to(dev)
    prob = _abi.make_problem(fittype, nc, nchan, npix, v[0], dv, lo, hi)
    outs[nc] = (prob, eng.fit(prob, rows, gd))
p1 = outs[1][0]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for what in ("aicc", "model2", "resid2"):
    ts = []
    for r in range(6):
        ev0.record()
        if what == "aicc":
            a = eng.aicc_global(p1, rows, outs[1][1]["params"], outs[2][1]["params"])
        elif what == "model2":
            m = eng.model(outs[2][0], outs[2][1]["params"])
        else:
            rs = eng.resid_stats(outs[2][0], rows, outs[2][1]["params"])
        ev1.record(); torch.cuda.synchronize(); ts.append(ev0.elapsed_time(ev1))
    print(os.environ.get("B200FIT_LIB", "current"), what, "ms: min %.3f median %.3f" % (min(ts), float(np.median(ts))))
print("ncomp hist", np.bincount(a["ncomp"].cpu().numpy().astype(int), minlength=3).tolist(), "lnk checksum %.9e" % float(torch.nan_to_num(a["lnk"]).sum()))
