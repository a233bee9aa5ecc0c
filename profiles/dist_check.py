#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): dist.fit_and_select on a cube partitioned by row slabs
must reproduce the single-GPU result bit for bit.   torchrun --nproc-per-node 2 profiles/dist_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mufasa_b200 import _abi, synth                      # noqa: E402
from mufasa_b200 import dist as D                         # noqa: E402
from mufasa_b200.engine import get_engine                 # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)


def gpu_modelcube(parcube, v, ft):
    P, ry, rx = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(eng.device)
    return eng.model(prob, par)[:, :len(v)].cpu().numpy().T.reshape(len(v), ry, rx)


ny = int(os.environ.get("DIST_NY", "61"))                 # odd: uneven slabs
syn = synth.make_cube(2, gpu_modelcube, rows=(100, 100 + ny), shape=(256, 96, 800))
t0 = time.perf_counter()
PART = os.environ.get("DIST_PARTITION", "cyclic")
res = D.fit_and_select(syn["cube"], syn["v"], syn["fittype"], syn["guess1"].copy(), syn["guess2"].copy(), device=local,
                       partition=PART)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if res is not None:
    res = {k: v.copy() for k, v in res.items()}            # views of a pinned buffer the next call reuses
if rank == 0:
    single = D.fit_and_select(syn["cube"], syn["v"], syn["fittype"], syn["guess1"].copy(), syn["guess2"].copy(),
                              device=local, group=dist.new_group([0]) if world > 1 else None)
else:
    if world > 1:
        dist.new_group([0])
if rank == 0:
    ok = True
    for k in res:
        same = np.array_equal(res[k], single[k], equal_nan=True)
        ok = ok and same
        print("%-13s %s  identical to single GPU: %s" % (k, res[k].shape, same))
    print("partition %s, world %d: partitioned job %.1f ms; ncomp hist %s; ALL IDENTICAL: %s" % (
        PART, world, dt * 1e3, np.bincount(res["ncomp"].astype(int).ravel(), minlength=3).tolist(), ok))
    assert ok
dist.barrier()
dist.destroy_process_group()
