"""Where the end-to-end time of dist.fit_and_select goes on ONE GPU at configs[4] size (1024 x 1024 x 1000):
wall-clock stamps around SelectionJob.__init__ (host work + queued uploads), run() (launches), the device
synchronisation and to_host, with CUDA events for the upload and the device work.  Usage: python ... [ny]"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mufasa_b200 import _abi, synth  # noqa: E402
from mufasa_b200 import dist as D  # noqa: E402
from mufasa_b200.engine import get_engine  # noqa: E402

CFG = 5
fittype, ny, nx, nchan, dv = synth.CONFIGS[CFG]
if len(sys.argv) > 1:
    ny = int(sys.argv[1])
eng = get_engine(0)
dev = eng.device


def gpu_modelcube(parcube, v, ft):
    P, py, px = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), py * px, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, py * px).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].float().cpu().numpy().T.reshape(len(v), py, px)


slab_t = torch.empty((nchan, ny, nx), dtype=torch.float32, pin_memory=True)
slab = slab_t.numpy()
g1, g2 = np.empty((4, ny, nx)), np.empty((8, ny, nx))
for a in range(0, ny, 64):
    b = min(ny, a + 64)
    part = synth.make_cube(CFG, gpu_modelcube, rows=(a, b), shape=(ny, nx, nchan), out=slab[:, a:b, :])
    g1[:, a:b], g2[:, a:b] = part["guess1"], part["guess2"]
    v = part["v"]
torch.cuda.empty_cache()


def once():
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    job = D.SelectionJob(slab, v, fittype, g1, g2, device=0)
    t.append(time.perf_counter())
    packed = job.run()
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    res = job.to_host(packed)
    t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])] + [1e3 * (t[-1] - t[0])]


for i in range(6):
    d = once()
    print("init %.1f  run(launch) %.1f  device wait %.1f  to_host %.1f  | total %.1f ms" % tuple(d))
import cProfile  # noqa: E402
import pstats  # noqa: E402

pr = cProfile.Profile()
pr.enable()
once()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
