import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
import _tiles as T
from mufasa_b200 import _abi, synth
from mufasa_b200.engine import get_engine
from oracle import model as M
eng = get_engine(0)
tile = T.sample_tile(2, nrows=2, xstep=16)
par = np.concatenate([tile["truth2"][:4]], axis=0)   # truth 1c params
par2 = tile["truth2"]
for nc, p in ((1, par), (2, par2)):
    P, ny, nx = p.shape
    ref = M.modelcube(tile["v"], p, tile["fittype"])
    prob = _abi.make_problem(tile["fittype"], nc, tile["nchan"], ny*nx, tile["v"][0], tile["dv"], [-1e30]*P, [1e30]*P)
    out = eng.model(prob, torch.from_numpy(np.ascontiguousarray(p.reshape(P,-1).T)).cuda()).cpu().numpy()[:, :tile["nchan"]].T.reshape(tile["nchan"], ny, nx)
    same = (out == ref)
    print("nc", nc, "bit-identical fraction %.5f" % same.mean(), " mask agree %.5f" % ((out>0)==(ref>0)).mean())
    d = np.abs(out-ref); nz = ~same
    # ulp distance for mismatches
    ulps = np.abs(out.view(np.int64) - ref.view(np.int64))[nz]
    print("   mismatches: n %d, median ulps %s, max ulps %s; |ref| at mismatch: median %.3g" % (nz.sum(), np.median(ulps), ulps.max(), np.median(np.abs(ref[nz]))))
    flips = ((out>0)!=(ref>0))
    print("   flips %d; ref values at flips:" % flips.sum(), np.sort(ref[flips])[:6], np.sort(out[flips])[-6:])
