"""Device-resident step of the selection job at 1024 x 1024 x 1000 on one GPU for several job-split limits."""
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
job = D.SelectionJob(slab, v, fittype, g1, g2, device=0)
job.run()
torch.cuda.synchronize()
for lim in [int(x) for x in os.environ.get("LIMITS", "2,3,4,6,8").split(",")]:
    os.environ["B200FIT_SPLIT_LIMIT"] = str(lim)
    job.run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for i in range(3):
        job.run()
        ev[i + 1].record()
    torch.cuda.synchronize()
    print("split limit %d: %s ms" % (lim, ["%.1f" % ev[i].elapsed_time(ev[i + 1]) for i in range(3)]))
