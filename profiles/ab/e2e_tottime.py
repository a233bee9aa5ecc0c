import sys, os, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mufasa_b200 import _abi, synth
from mufasa_b200.cube import SpecCube
from mufasa_b200.engine import get_engine
from mufasa_b200.UltraCube import UltraCube
eng = get_engine(0); dev = eng.device
CFG = 2
fittype, ny, nx, nchan, dv = synth.CONFIGS[CFG]
def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].cpu().numpy().T.reshape(len(v), ry, rx)
pinned = torch.empty((nchan, ny, nx), dtype=torch.float32).pin_memory()
syn = synth.make_cube(CFG, gpu_modelcube, out=pinned.numpy())
spec = SpecCube(syn["cube"], syn["v"])
def step():
    uc = UltraCube(cube=spec, fittype=fittype, device=0)
    uc.fit_cube(ncomp=[1, 2], simpfit=True, guesses={1: syn["guess1"].copy(), 2: syn["guess2"].copy()})
    res = uc.get_best_2c_parcube()
    torch.cuda.synchronize()
    return uc
for _ in range(3): step()
ts=[]
for _ in range(5):
    t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
print("e2e steps ms", ["%.1f" % t for t in ts], "slabs", len((step()._ref_pcube._dev.get("slab_ready") or ([None], 0))[0]))
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
