#!/usr/bin/env python
"""bench.py -- throughput of the B200 per-pixel spectral fitter on BASELINE.json's metric.

metric   : spectra fitted / second, 2-component NH3(1,1) model-selection workload
workload : BASELINE.json configs[4] (the configuration the north-star target is quoted on) -- ONE synthetic NH3(1,1)
           cube, 1024 x 1024 pixels x 1000 channels.  One STEP = the whole job on the whole cube: 1-component fit +
           2-component fit of every spectrum + AICc 0/1/2-component selection + assembly of the best-model parameter
           cube (UltraCube.fit_cube([1, 2]) + get_best_2c_parcube in the reference's terms).
           N GPUs: the SAME cube partitioned by pixel blocks (mufasa_b200.dist.partition_row_blocks: blocks of 8 rows
           dealt round-robin, so every rank gets the same mix of easy and hard spectra) -> "scaling": "strong".  Pixels are independent: no collective inside the fit
           loop; NCCL carries the global arg-max of the AICc include_nosamp rule (16 bytes per rank + one 256-byte
           broadcast) and ONE gather of the [44, rows, nx] result block per step.
value    : 1 048 576 spectra / (max-over-ranks device time per step), slab rows and guesses already resident in HBM;
           the step runs from the launch of the two fits to the gathered result block on rank 0's GPU.
e2e      : the same job through the public host API, mufasa_b200.dist.fit_and_select, starting from HOST (pinned)
           buffers: H2D of each rank's slab and guesses, all kernels, the NCCL gather, D2H of the 44 result maps on
           rank 0; wall clock between barriers, max over ranks.
roofline : dominant kernel = fit_eval_kernel<2> (model + analytic Jacobian + normal equations of the 2-component
           fit).  The path is compute bound.  ``frac`` is the EXECUTED fraction of the binding pipe
           (XU: ex2 actually issued / kernel time / the peak b200fit_measure_xu_peak measures on this GPU); the algorithmic figure (what the reference's arithmetic
           would need, most of which the kernel's windows skip) is reported beside it as ``frac_algorithmic``.
cpu_baseline / --impl reference: the CPU oracle (numpy restatement of the reference's pyspeckit/mpfit path; the
           reference itself is pure Python on un-vendored pyspeckit and cannot be installed here) timed on this
           box's host cores on ONE fixed sample of the same cube (REF_ROWS x every REF_XSTEP-th column).

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = 5                                  # synth.CONFIGS key of BASELINE.json configs[4]
METRIC = "spectra fitted/sec, 2-comp NH3(1,1)"
UNIT = "spectra/s"
REF_NROWS, REF_XSTEP = 8, 32             # the CPU sample: 8 rows x every 32nd column of the cube = 256 spectra


# ------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler(object):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a thread
    (nvidia_ml_py); a polling `nvidia-smi -lms` process takes the driver lock for milliseconds per query and slows a
    launch-latency-bound loop like this one by ~30 %, so it is only the fallback."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.05):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.thread, self.stop_flag, self.nvml, self.smax = None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, pw))
                for name, bit in self.BAD.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self.nvml is None:
            return self._smi_once()
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.smax, reasons=["no samples"])
        sm = [x[0] for x in self.samples]
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=self.smax, power_w_max=float(max(x[1] for x in self.samples)),
                    samples=len(sm), reasons=sorted(self.reasons), how="NVML polled every %d ms" % int(self.period * 1e3))

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [float(x) for x in out.strip().split(",")]
            return dict(sm_mhz=f[0], sm_max_mhz=f[1], reasons=[], how="one nvidia-smi query after the timed region")
        except Exception:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml and nvidia-smi unavailable"])


# ------------------------------------------------------------------------------------------------------------------
def oracle_modelcube(parcube, v, fittype):
    """Model cube for synth.make_cube from the CPU oracle -- only at the sampled columns (the generator draws the
    noise of whole rows, so the sampled spectra are those of the full cube; the other columns are not used)."""
    from oracle import model as M
    out = np.zeros((len(v),) + parcube.shape[1:])
    out[:, :, ::REF_XSTEP] = M.modelcube(v, parcube[:, :, ::REF_XSTEP], fittype)
    return out


def sample_rows(ny, nrows):
    return [int(y) for y in np.linspace(20, ny - 21, nrows)]


def cpu_sample(cfg):
    """The fixed CPU sample of the workload cube (same cube, same guesses as the GPU arm: synth.make_cube is
    reproducible per row): dict(cube, g1, g2, v, fittype, vstats-free limits)."""
    from mufasa_b200 import synth
    fittype, ny, nx, nchan, dv = synth.CONFIGS[cfg]
    parts = [synth.make_cube(cfg, oracle_modelcube, rows=(y, y + 1)) for y in sample_rows(ny, REF_NROWS)]
    cube = np.ascontiguousarray(np.concatenate([p["cube"][:, :, ::REF_XSTEP] for p in parts], axis=1))
    g1 = np.concatenate([p["guess1"][:, :, ::REF_XSTEP] for p in parts], axis=1)
    g2 = np.concatenate([p["guess2"][:, :, ::REF_XSTEP] for p in parts], axis=1)
    return dict(cube=cube, g1=g1, g2=g2, v=parts[0]["v"], fittype=fittype,
                what="%d spectra (%d rows x every %d-th column of the %dx%dx%d cube)" % (
                    cube.shape[1] * cube.shape[2], REF_NROWS, REF_XSTEP, ny, nx, nchan))


def cpu_job(sample, cores):
    """The reference's job on the sample, on the host: 1c fit + 2c fit (cubefit_simp semantics: guesses clamped into
    the limits, err = std(spectrum), mpfit) + get_all_lnk_maps / get_best_2c_parcube.  Returns (nspectra, seconds)."""
    from mufasa_b200 import synth
    from oracle import constants as K
    from oracle import fiteach as F
    from oracle import model as M
    from oracle import stats as ST
    cube, v, fittype = sample["cube"], sample["v"], sample["fittype"]
    L = K.limits(fittype)
    lo1 = [-3.5, L["sigmin"], L["Texmin"], L["taumin"]]
    hi1 = [4.5, L["sigmax"], L["Texmax"], L["taumax"]]
    g1 = synth.clamp_guesses(sample["g1"], lo1, hi1)
    g2 = synth.clamp_guesses(sample["g2"], lo1 * 2, hi1 * 2)
    nspec = cube.shape[1] * cube.shape[2]
    assert nspec >= 2 * cores, "oracle.fiteach runs serially below 2 spectra per core"
    t0 = time.perf_counter()
    o1 = F.fiteach(cube, v, g1, lo1, hi1, fittype, multicore=cores)
    o2 = F.fiteach(cube, v, g2, lo1 * 2, hi1 * 2, fittype, multicore=cores)
    m1 = M.modelcube(v, o1["parcube"], fittype, dtype=np.float32)
    m2 = M.modelcube(v, o2["parcube"], fittype, dtype=np.float32)
    with np.errstate(all="ignore"):
        Lk = ST.all_lnk_maps(cube, m1, m2, expand=20)
        ST.best_2c_parcube(o1["parcube"], o1["errcube"], o2["parcube"], o2["errcube"], Lk["lnk10"], Lk["lnk20"],
                           Lk["lnk21"])
    return nspec, time.perf_counter() - t0


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle restatement) on this box's host cores, all of them, on the
    fixed sample -- the same sample at every N and in the b200 arm's ``cpu_baseline``."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = min(host_cores(), REF_NROWS * (1024 // REF_XSTEP) // 4)     # >= 4 spectra per worker process
    sample = cpu_sample(CFG)
    times, nspec = [], 0
    t_begin = time.perf_counter()
    done_warm = 0
    for it in range(args.warmup + args.steps):
        # CPU runs have no warm-up effect beyond the first repetition (fork + page cache): when the whole run would
        # not fit in ~4 minutes the remaining warm-up repetitions are skipped, never the timed ones
        if it < args.warmup and it >= 1 and times_est(t_begin, it) * (args.steps + args.warmup) > 240.0:
            continue
        nspec, dt = cpu_job(sample, cores)
        if it >= args.warmup:
            times.append(dt)
        else:
            done_warm += 1
    tstep = float(np.mean(times))
    value = nspec / tstep
    what = "%s per step; 1c fit + 2c fit + AICc; %d worker processes; %d of %d warm-up repetitions run" % (
        sample["what"], cores, done_warm, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tstep * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = pure-Python MUFASA on un-vendored pyspeckit/mpfit (not installable offline); timed "
                    "here: the numpy oracle restatement of that path (oracle/), multiprocessing over the host cores"}
    print(json.dumps(line))
    return 0


def times_est(t_begin, reps_done):
    return (time.perf_counter() - t_begin) / max(1, reps_done)


def workload_config(n):
    return {"workload": "BASELINE configs[4]: ONE synthetic NH3(1,1) cube 1024x1024x1000 ch; step = 1-comp fit + "
                        "2-comp fit + AICc ncomp selection + best-model cube of all 1048576 spectra",
            "spectra": 1048576, "nchan": 1000, "fittype": "nh3_multi_v",
            "partition": ("blocks of 8 rows of the one cube dealt round-robin to %d GPUs (%d rows each), result block "
                          "gathered to rank 0 over NCCL" % (n, 1024 // n)) if n > 1 else "single GPU, whole cube",
            "cache": "inputs larger than L2 (%.1f GB of spectra per GPU vs 126 MB L2), streamed from HBM by every "
                     "kernel" % (4.29 / max(1, n))}


# ------------------------------------------------------------------------------------------------------------------
class RankRows(object):
    """[nchan, ny, nx] stand-in that only holds this rank's rows (stacked, in pinned memory) -- what a loader that
    reads its own rows hands fit_and_select (mufasa_b200.dist.take_rows)."""

    def __init__(self, slab, ny, rows):
        self.slab, self.rows = slab, np.asarray(rows)
        self.shape = (slab.shape[0], ny, slab.shape[2])

    def take_rows(self, rows):
        assert np.array_equal(np.asarray(rows), self.rows)
        return self.slab


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mufasa_b200 import _abi, synth
    from mufasa_b200 import dist as D
    from mufasa_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine(local)
    dev = eng.device
    warmup = max(3, args.warmup)

    fittype, ny, nx, nchan, dv = synth.CONFIGS[CFG]
    if args.small:                              # smoke-size run of the same code path (not a bench line)
        ny, nx = 64 * world, 128
    nspec = ny * nx
    my_rows = D.owned_rows(ny, world, rank, "cyclic")      # blocks of 8 rows dealt round-robin: even load per rank
    ry = int(my_rows.size)

    def gpu_modelcube(parcube, v, ft):
        P, py, px = parcube.shape
        prob = _abi.make_problem(ft, P // 4, len(v), py * px, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
        par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, py * px).T)).to(dev)
        return eng.model(prob, par)[:, :len(v)].float().cpu().numpy().T.reshape(len(v), py, px)

    # ---------------- synthesis (untimed): this rank's slab into pinned memory, 64 rows at a time ------------------
    t_syn = time.perf_counter()
    slab_t = torch.empty((nchan, ry, nx), dtype=torch.float32, pin_memory=True)
    slab = slab_t.numpy()
    g1s, g2s = np.empty((4, ry, nx)), np.empty((8, ry, nx))
    shape = (ny, nx, nchan) if args.small else None
    a = 0
    while a < ry:                               # runs of consecutive rows (whole cube at N = 1: 64 rows at a time)
        b = a + 1
        while b < ry and my_rows[b] == my_rows[b - 1] + 1 and b - a < 64:
            b += 1
        part = synth.make_cube(CFG, gpu_modelcube, rows=(int(my_rows[a]), int(my_rows[b - 1]) + 1), shape=shape,
                               out=slab[:, a:b, :])
        g1s[:, a:b], g2s[:, a:b] = part["guess1"], part["guess2"]
        v = part["v"]
        a = b
    torch.cuda.empty_cache()

    def all_rows(a):                            # [k, ry, nx] maps of every rank's rows -> [k, ny, nx] on every rank
        if world == 1:
            return a
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        pieces = [torch.empty((a.shape[0], D.owned_rows(ny, world, r, "cyclic").size, nx), dtype=t.dtype, device=dev)
                  for r in range(world)]
        dist.all_gather(pieces, t)
        out = np.empty((a.shape[0], ny, nx), dtype=a.dtype)
        for r, piece in enumerate(pieces):
            out[:, D.owned_rows(ny, world, r, "cyclic"), :] = piece.cpu().numpy()
        return out

    g1_full, g2_full = all_rows(g1s), all_rows(g2s)
    t_syn = time.perf_counter() - t_syn
    cube_view = RankRows(slab, ny, my_rows)

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---------------- end-to-end arm: the public API from host buffers ------------------------------------------------
    e2e_times = []
    n_e2e = max(2, min(args.steps, 5))
    res = None
    for it in range(2 + n_e2e):
        barrier()
        t0 = time.perf_counter()
        res = D.fit_and_select(cube_view, v, fittype, g1_full, g2_full, device=local)
        barrier()
        if it >= 2:
            e2e_times.append(max_over_ranks(time.perf_counter() - t0))
    e2e_s = float(np.mean(e2e_times))
    e2e_value = nspec / e2e_s
    ncomp_hist_e2e = np.bincount(res["ncomp"].astype(int).ravel(), minlength=3).tolist() if rank == 0 else None
    d2h = int(_abi.PACK_FIELDS * nspec * 8) + 16 * 2

    # ---------------- device arm: slab rows and guesses resident in HBM -------------------------------------------------
    job = D.SelectionJob(cube_view, v, fittype, g1_full, g2_full, device=local)
    h2d_t = torch.tensor([float(job.h2d_bytes())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_t)
    h2d = int(h2d_t.item())
    for _ in range(warmup):
        packed = job.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count()
    host_launches0 = eng.host_launch_count()
    start.record()
    for _ in range(args.steps):
        packed = job.run()
    stop.record()
    barrier()
    launches_timed = eng.launch_count() - launches0
    host_launches_timed = eng.host_launch_count() - host_launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = max_over_ranks(start.elapsed_time(stop)) / args.steps
    value = nspec / (ms_per_step * 1e-3)
    ncomp_hist = None
    if rank == 0:
        ncomp_hist = torch.bincount(packed[43].to(torch.int64).ravel(), minlength=3).tolist()
    st2 = job.fits[2]["status"]
    conv = torch.tensor([float((((st2 >= 1) & (st2 <= 4)) | (st2 == 6)).sum()), float((st2 == 5).sum()),
                         float(job.fits[2]["counts"][:, 0].sum()), float(job.fits[1]["counts"][:, 0].sum())],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(conv)

    # per-stage device times of one more step on this rank (fits | selection + assembly + gather), max over ranks
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    barrier()
    ev[0].record()
    outs = eng.fit_batch(job.batch)
    ev[1].record()
    sel = D.aicc_distributed(eng, job.prob, job.ref.rows_on_device(), outs[job.jobs[1]]["params"],
                             outs[job.jobs[2]]["params"], model_f32=True,
                             local_to_global=lambda i: int(my_rows[i // nx]) * nx + i % nx)
    eng.pack_selection(outs[job.jobs[1]], outs[job.jobs[2]], sel)
    ev[2].record()
    barrier()
    t_fits = max_over_ranks(ev[0].elapsed_time(ev[1]))
    t_sel = max_over_ranks(ev[1].elapsed_time(ev[2]))
    del outs, sel

    # ---------------- roofline of the dominant kernel (fit_eval_kernel<2>) on rank 0's slab -------------------------
    # One extra fit of each kind, outside the timed region, with CUDA events recorded by the library around every
    # kernel on the launching stream (b200fit_set_profiling); a profiled fit runs undivided.
    peaks = measured_peaks()
    j1, j2 = job.batch[job.jobs[1]], job.batch[job.jobs[2]]
    eng.set_profiling(True)
    o1 = eng.fit(j1["prob"], j1["spectra"], j1["guesses"], row_index=j1["row_index"])
    prof1 = eng.get_profile()
    o2 = eng.fit(j2["prob"], j2["spectra"], j2["guesses"], row_index=j2["row_index"])
    prof2 = eng.get_profile()
    eng.set_profiling(False)
    npix_r = int(j2["prob"].npix)
    cnt = o2["counts"].to(torch.float64)
    evals = float(cnt[:, 0].sum().item())                  # fused model + Jacobian evaluations, summed over pixels
    exec_gauss = float(cnt[:, 2].sum().item()) * 32.0       # gaussians actually evaluated (windowed)
    exec_chan = float(cnt[:, 3].sum().item()) * 32.0        # channels actually visited (active 32-channel chunks)
    H, C, P = 18, 2, 8
    sfu = evals * nchan * C * (H + 2)                       # SFU_J = N*C*(H+2)
    f32 = evals * nchan * C * (12 * H + 30)                 # F32_J = N*C*(12H+30)
    f64 = evals * nchan * (P * (P + 1) + 2 * P + 2)         # F64_J = N*(P(P+1)+2P+2)
    sfu_exec = exec_gauss + exec_chan * C                   # ex2 actually issued: gaussians + e^-tau per component
    f32_exec = exec_gauss * 8.0 + exec_chan * (40.0 + 2.0 * 45.0)   # per gaussian 8, per channel RT 40 + 45 FMA
    launches_eval = max(1, prof2["rounds"])
    t_eval = prof2["eval_ms"] * 1e-3                        # sum over the launches of one fit
    clk = peaks["sm_max_mhz"] * 1e6
    nsm = eng.num_sms
    xu_measured = eng.measure_xu_peak()                    # in-repo micro-benchmark on this box
    pk = {"fp32": nsm * 128 * 2 * clk, "xu": xu_measured, "fp64": nsm * 64 * 2 * clk}
    ach = {"fp32": f32 / t_eval, "xu": sfu / t_eval, "fp64": f64 / t_eval}
    frac_alg = {k: ach[k] / pk[k] for k in pk}
    frac_exec = {"xu": sfu_exec / t_eval / pk["xu"], "fp32": f32_exec / t_eval / pk["fp32"]}
    t2 = (prof2["prologue_ms"] + prof2["eval_ms"] + prof2["solve_ms"]) * 1e-3   # one 2c fit of this slab on its own
    hbm_bytes = npix_r * (nchan * 4 + P * 8 + (2 * P + 2) * 8 + 4 + 16)
    row_bytes = (nchan + 31) // 32 * 32 * 4
    traffic = ncu_traffic_per_eval(row_bytes) * evals / launches_eval
    roofline = {"bound": "xu", "kernel": "fit_eval_kernel<2>", "achieved": sfu_exec / t_eval / 1e12,
                "peak": pk["xu"] / 1e12, "unit": "Top/s", "frac": frac_exec["xu"],
                "definition": "EXECUTED ex2 (gaussians inside the 6.2-sigma windows + one e^-tau per channel and "
                              "component, counted by the kernel) / summed CUDA-event duration of the kernel's launches "
                              "in one undivided 2c fit of rank 0's slab / XU peak",
                "frac_algorithmic": frac_alg, "frac_executed": frac_exec,
                "algorithmic_definition": "SURVEY 8d per-evaluation figures x evaluations counted by the kernel; the "
                                          "kernel skips gaussians beyond 6.2 sigma (exactly 0 in fp32), ~90 % of them",
                "traffic": traffic,
                "traffic_source": "dram bytes per pixel-evaluation from the latest ncu --set full capture under "
                                  "profiles/ (NCU_TRAFFIC.json: %.0f B at this row length; algorithmic %d B) x "
                                  "evaluations per launch of this run" % (ncu_traffic_per_eval(row_bytes),
                                                                          row_bytes + 112 + 360 + 4),
                "peak_source": "measured: b200fit_measure_xu_peak micro-benchmark on this GPU (8 independent ex2.approx "
                               "chains per thread); derived %d SMs x 16 XU lanes x %.0f MHz = %.3f Top/s" % (
                                   nsm, peaks["sm_max_mhz"], nsm * 16 * clk / 1e12),
                "hbm": {"achieved": hbm_bytes / t2 / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": hbm_bytes / t2 / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["source"]},
                "per_launch": {"launches": launches_eval, "avg_ms": prof2["eval_ms"] / launches_eval,
                               "xu_ops_executed": sfu_exec / launches_eval, "xu_ops_algorithmic": sfu / launches_eval},
                "per_fit": {"pixels": npix_r, "evaluations": evals, "executed_gaussians": exec_gauss,
                            "algorithmic_gaussians": evals * nchan * C * H},
                "fit_2c_ms": {"prologue": prof2["prologue_ms"], "eval": prof2["eval_ms"], "solve": prof2["solve_ms"],
                              "rounds": prof2["rounds"]},
                "fit_1c_ms": {"prologue": prof1["prologue_ms"], "eval": prof1["eval_ms"], "solve": prof1["solve_ms"],
                              "rounds": prof1["rounds"]},
                "share_of_step": {"fits_1c_and_2c_batched": t_fits / (t_fits + t_sel),
                                  "selection_assembly_gather": t_sel / (t_fits + t_sel)}}
    del o1, o2

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 model+Jacobian, f64 normal equations/solve",
                "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                        "samples_ms": [round(t * 1e3, 2) for t in e2e_times],
                        "api": "mufasa_b200.dist.fit_and_select(cube slab view, v, fittype, guesses1, guesses2): "
                               "pinned host slab in, 44 gathered maps out on rank 0; wall clock, max over ranks"},
                "gpu_launches": int(launches_timed),
                "host_launches": int(host_launches_timed),   # kernels outside graphs + graph launches (8 rounds each)
                "stage_ms": {"fits_1c_2c": t_fits, "selection_assembly_gather": t_sel},
                "roofline": roofline, "clocks": clocks,
                "fit_stats": {"converged_2c": conv[0].item() / nspec, "maxiter_2c": conv[1].item() / nspec,
                              "mean_evals_2c": conv[2].item() / nspec, "mean_evals_1c": conv[3].item() / nspec,
                              "ncomp_hist": ncomp_hist, "ncomp_hist_e2e": ncomp_hist_e2e},
                "synthesis_s": round(t_syn, 1)}
        if args.small:
            line["config"]["workload"] = "SMOKE-SIZE run (%dx%dx%d), not a bench line" % (ny, nx, nchan)
    barrier()
    # ---------------- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(host_cores(), REF_NROWS * (1024 // REF_XSTEP) // 4)
        sample = cpu_sample(CFG)
        ns, dt = cpu_job(sample, cores)
        line["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%s, 1c fit + 2c fit + AICc with the numpy oracle (mpfit restatement), %d "
                                          "worker processes, %.1f s" % (sample["what"], cores, dt)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic_per_eval(row_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum per pixel-evaluation of fit_eval_kernel<2>, from the latest
    ncu --set full capture of a launch with every pixel active (profiles/NCU_TRAFFIC.json, written next to the
    capture's summary): the part that does not depend on the channel count (per-pixel inputs, normal equations out)
    as measured, plus the spectrum row of THIS workload (the kernel stages the whole row once per evaluation; the
    capture's own row accounted for its 800 channels to within 5 %)."""
    path = os.path.join(ROOT, "profiles", "NCU_TRAFFIC.json")
    with open(path) as f:
        return float(json.load(f)["dram_bytes_per_pixel_eval_minus_row"]) + float(row_bytes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostics only")
    ap.add_argument("--small", action="store_true", help="smoke-size cube through the same code path (diagnostics)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
