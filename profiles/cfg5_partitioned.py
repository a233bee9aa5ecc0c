#!/usr/bin/env python
"""BASELINE configs[4]: the synthetic NH3(1,1) 1024x1024x1000 cube, pixel-partitioned by row slabs over the GPUs of
one box (strong scaling: the cube is fixed, each rank holds ny / world rows).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/cfg5_partitioned.py

Each rank synthesises only its own slab (the whole cube is 4 GB); the guess maps are all-gathered because
``dist.fit_and_select`` wants the global velocity statistics of the whole guess map.  Reported, max over ranks:
  * device-timed (CUDA events, rows resident): the 2-component fit alone, and the 1c || 2c batch + AICc selection;
  * the public API end to end: ``dist.fit_and_select`` from host slabs to the gathered maps on rank 0 (wall clock
    between barriers, H2D and D2H and the NCCL gather inside)."""
import json
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
from mufasa_b200.spec_models.meta_model import MetaModel  # noqa: E402

CFG = int(os.environ.get("PART_CFG", "5"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
dev = eng.device
fittype, ny, nx, nchan, dv = synth.CONFIGS[CFG]
assert ny % world == 0, "even slabs keep the all-gather of the guess maps simple"
y0, y1 = D.partition_rows(ny, world, rank)
ry = y1 - y0


def gpu_modelcube(parcube, v, ft):
    P, py, px = parcube.shape
    prob = _abi.make_problem(ft, P // 4, len(v), py * px, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
    par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, py * px).T)).to(dev)
    return eng.model(prob, par)[:, :len(v)].float().cpu().numpy().T.reshape(len(v), py, px)


class RowSlabView(object):
    """[nchan, ny, nx] stand-in that only holds rows y0:y1 (what a slab-reading loader would hand over)."""

    def __init__(self, slab, ny, y0):
        self.slab, self.y0 = slab, y0
        self.shape = (slab.shape[0], ny, slab.shape[2])

    def __getitem__(self, key):
        rows = key[1]
        assert key[0] == slice(None) and (rows.start, rows.stop) == (self.y0, self.y0 + self.slab.shape[1])
        return self.slab


def all_rows(a):
    """[k, ry, nx] slab maps of every rank -> [k, ny, nx] on every rank."""
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, t)
    return out.permute(1, 0, 2, 3).reshape(t.shape[0], ny, nx).cpu().numpy()


t0 = time.perf_counter()
PINNED = os.environ.get("PART_PINNED", "1") == "1"    # the slab in pinned memory: slab-pipelined, asynchronous H2D
slab_host = torch.empty((nchan, ry, nx), dtype=torch.float32, pin_memory=True).numpy() if PINNED else None
syn = synth.make_cube(CFG, gpu_modelcube, rows=(y0, y1), out=slab_host)
v = syn["v"]
g1_full, g2_full = all_rows(syn["guess1"]), all_rows(syn["guess2"])
t_syn = time.perf_counter() - t0


def max_over_ranks(x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- device-timed, rows resident -------------------------------------------------------------------------------
npix = ry * nx
mm = MetaModel(fittype, 1)
cube_dev = torch.from_numpy(syn["cube"].reshape(nchan, npix)).to(dev)
rows = eng.gather(cube_dev)
del cube_dev
probs, gs = {}, {}
for nc in (1, 2):
    lo = [-3.5, mm.sigmin, mm.Texmin, mm.taumin] * nc
    hi = [4.5, mm.sigmax, mm.Texmax, mm.taumax] * nc
    g = synth.clamp_guesses(syn["guess%d" % nc], lo, hi, eps=mm.eps)
    gs[nc] = torch.from_numpy(np.ascontiguousarray(g.reshape(4 * nc, npix).T)).to(dev)
    probs[nc] = _abi.make_problem(fittype, nc, nchan, npix, v[0], dv, lo, hi)


def timed(fn, reps=3):
    best = []
    for _ in range(reps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(max_over_ranks(e0.elapsed_time(e1)))
    return best, out


def step():
    o = eng.fit_batch([dict(prob=probs[1], spectra=rows, guesses=gs[1]),
                       dict(prob=probs[2], spectra=rows, guesses=gs[2])])
    return o, eng.aicc_global(probs[1], rows, o[0]["params"], o[1]["params"])


fit2_ms, o2 = timed(lambda: eng.fit(probs[2], rows, gs[2]))
step_ms, (ob, sel) = timed(step)
hist = torch.bincount(sel["ncomp"].to(torch.int64), minlength=3).to(torch.float64)
dist.all_reduce(hist)
st = o2["status"]
conv = torch.tensor([float(((st >= 1) & (st <= 4) | (st == 6)).sum()), float((st == 5).sum())], dtype=torch.float64,
                    device=dev)
dist.all_reduce(conv)
del rows, o2, ob, sel
torch.cuda.empty_cache()

# ---- the public API, host slabs in, gathered maps out ----------------------------------------------------------
vcube = RowSlabView(syn["cube"], ny, y0)
e2e_ms = []
for _ in range(3):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    res = D.fit_and_select(vcube, v, fittype, g1_full, g2_full, device=local)
    torch.cuda.synchronize()
    dist.barrier()
    e2e_ms.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
if rank == 0:
    nspec = ny * nx
    line = {
        "workload": "BASELINE configs[4]: synthetic NH3(1,1) %dx%dx%d ch, row slabs over %d GPU(s) (strong scaling)"
                    % (ny, nx, nchan, world),
        "n_gpus": world, "spectra": nspec, "host_slab": "pinned" if PINNED else "pageable", "rows_per_gpu": ry, "synthesis_s_per_rank": round(t_syn, 1),
        "fit_2c_ms": [round(x, 2) for x in fit2_ms],
        "fit_2c_spectra_per_s": nspec / (min(fit2_ms) * 1e-3),
        "step_1c_2c_aicc_ms": [round(x, 2) for x in step_ms],
        "step_spectra_per_s": nspec / (min(step_ms) * 1e-3),
        "e2e_fit_and_select_ms": [round(x, 1) for x in e2e_ms],
        "e2e_spectra_per_s": nspec / (min(e2e_ms) * 1e-3),
        "fit_2c_converged_frac": conv[0].item() / nspec, "fit_2c_maxiter_frac": conv[1].item() / nspec,
        "ncomp_hist_selected": [int(x) for x in hist.tolist()],
        "ncomp_hist_e2e": np.bincount(res["ncomp"].astype(int).ravel(), minlength=3).tolist(),
        "timing": "device: CUDA events between barriers, max over ranks, best of 3 (all listed); e2e: wall clock "
                  "between barriers, max over ranks",
    }
    print(json.dumps(line), flush=True)
dist.barrier()
dist.destroy_process_group()
