#!/usr/bin/env python
"""bench.py -- throughput of the B200 per-pixel spectral fitter on BASELINE.json's metric.

metric   : spectra fitted / second, 2-component NH3(1,1) model selection workload
workload : BASELINE.json configs[1] -- synthetic NH3(1,1) cube 256 x 256 x 800 channels; one STEP = the whole job
           on one cube: 1-component fit + 2-component fit of every spectrum + AICc 0/1/2-component selection
           (UltraCube.fit_cube(1), fit_cube(2), get_best_2c_parcube in the reference's terms).
           N GPUs: every rank owns one such cube (independent noise realisation) -- pixels are independent, no
           collective on the data path -> "scaling": "weak".
value    : (spectra of all ranks) / (max-over-ranks device time per step), cube rows already resident in HBM.
e2e      : the same job through the public host API (mufasa_b200.UltraCube) starting from HOST (pinned) buffers:
           H2D of the cube and guesses, all kernels, D2H of parameter / error / lnk / ncomp maps, host bookkeeping.
roofline : dominant kernel = fit_eval_kernel<2> (model + Jacobian + normal equations of the 2-component fit).  The path is compute bound: the
           fraction reported is algorithmic pipe work / time / pipe peak for the binding pipe (FP32 FMA, XU ex2,
           FP64), with the HBM figure alongside.
cpu_baseline / --impl reference: the CPU oracle (numpy restatement of the reference's pyspeckit/mpfit path; the
           reference itself is pure Python on un-vendored pyspeckit and cannot be installed here) timed on this
           box's host cores on a bounded sample of the same cube.

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

CFG = 2
METRIC = "spectra fitted/sec, 2-comp NH3(1,1)"
UNIT = "spectra/s"


# ------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


NCU_DRAM_BYTES_PER_EVAL = (166.019072e6 + 19.493888e6) / 65536.0   # profiles/r1m_fit2c_ncu_full_summary.txt


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
    from oracle import model as M
    return M.modelcube(v, parcube, fittype)


def sample_rows(ny, nrows):
    return [int(y) for y in np.linspace(20, ny - 21, nrows)]


def cpu_job(cfg, rows, xstep, cores, seed_offset=0):
    """The reference's job on a sample of the workload cube, on the host: 1c fit + 2c fit (cubefit_simp semantics:
    guesses clamped into the limits, err = std(spectrum), mpfit) + get_all_lnk_maps / get_best_2c_parcube.
    Returns (nspectra, seconds)."""
    from mufasa_b200 import synth
    from oracle import constants as K
    from oracle import fiteach as F
    from oracle import model as M
    from oracle import stats as ST
    parts = [synth.make_cube(cfg, oracle_modelcube, rows=(y, y + 1), seed_offset=seed_offset) for y in rows]
    cube = np.ascontiguousarray(np.concatenate([p["cube"][:, :, ::xstep] for p in parts], axis=1))
    g1 = np.concatenate([p["guess1"][:, :, ::xstep] for p in parts], axis=1)
    g2 = np.concatenate([p["guess2"][:, :, ::xstep] for p in parts], axis=1)
    v, fittype = parts[0]["v"], parts[0]["fittype"]
    L = K.limits(fittype)
    lo1 = [-3.5, L["sigmin"], L["Texmin"], L["taumin"]]
    hi1 = [4.5, L["sigmax"], L["Texmax"], L["taumax"]]
    g1 = synth.clamp_guesses(g1, lo1, hi1)
    g2 = synth.clamp_guesses(g2, lo1 * 2, hi1 * 2)
    t0 = time.perf_counter()
    o1 = F.fiteach(cube, v, g1, lo1, hi1, fittype, multicore=cores)
    o2 = F.fiteach(cube, v, g2, lo1 * 2, hi1 * 2, fittype, multicore=cores)
    m1 = M.modelcube(v, o1["parcube"], fittype, dtype=np.float32)
    m2 = M.modelcube(v, o2["parcube"], fittype, dtype=np.float32)
    with np.errstate(all="ignore"):
        Lk = ST.all_lnk_maps(cube, m1, m2, expand=20)
        ST.best_2c_parcube(o1["parcube"], o1["errcube"], o2["parcube"], o2["errcube"], Lk["lnk10"], Lk["lnk20"],
                           Lk["lnk21"])
    dt = time.perf_counter() - t0
    return cube.shape[1] * cube.shape[2], dt


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle restatement) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from mufasa_b200 import synth
    fittype, ny, nx, nchan, dv = synth.CONFIGS[CFG]
    cores = host_cores()
    # pilot: one short row sample to size the per-step sample so that the whole run stays within ~3 minutes
    n0, t0 = cpu_job(CFG, sample_rows(ny, 2), 16, cores)
    rate = n0 / t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    want = int(min(2048, max(32, rate * budget)))
    nrows = max(2, min(16, want // 32))
    xstep = max(1, nx // max(1, want // nrows))
    rows = sample_rows(ny, nrows)
    times = []
    nspec = 0
    for it in range(args.warmup + args.steps):
        nspec, dt = cpu_job(CFG, rows, xstep, cores)
        if it >= args.warmup:
            times.append(dt)
    tstep = float(np.mean(times))
    value = nspec / tstep
    sample = "%d spectra per step (%d rows x every %d-th column of the %dx%dx%d cube); 1c fit + 2c fit + AICc" % (
        nspec, nrows, xstep, ny, nx, nchan)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tstep * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = pure-Python MUFASA on un-vendored pyspeckit/mpfit (not installable offline); timed "
                    "here: the numpy oracle restatement of that path (oracle/), multiprocessing over all host cores"}
    print(json.dumps(line))
    return 0


def workload_config(n):
    return {"workload": "BASELINE configs[1]: synthetic NH3(1,1) cube 256x256x800 ch per GPU; step = 1-comp fit + "
                        "2-comp fit + AICc ncomp selection of all 65536 spectra",
            "spectra_per_gpu": 65536, "nchan": 800, "fittype": "nh3_multi_v",
            "partition": "one cube per rank (pixels independent, no data-path collective)" if n > 1 else "single GPU",
            "cache": "inputs larger than L2 (210 MB of spectra per GPU vs 126 MB L2), streamed once per kernel"}


# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from mufasa_b200 import _abi, synth
    from mufasa_b200.cube import SpecCube
    from mufasa_b200.engine import get_engine
    from mufasa_b200.spec_models.meta_model import MetaModel
    from mufasa_b200.UltraCube import UltraCube

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

    fittype, ny, nx, nchan, dv = synth.CONFIGS[CFG]
    npix = ny * nx

    def gpu_modelcube(parcube, v, ft):
        P, ry, rx = parcube.shape
        prob = _abi.make_problem(ft, P // 4, len(v), ry * rx, v[0], v[1] - v[0], [-1e30] * P, [1e30] * P)
        par = torch.from_numpy(np.ascontiguousarray(parcube.reshape(P, ry * rx).T)).to(dev)
        m = eng.model(prob, par)[:, :len(v)]
        return m.cpu().numpy().T.reshape(len(v), ry, rx)

    pinned = torch.empty((nchan, ny, nx), dtype=torch.float32).pin_memory()
    syn = synth.make_cube(CFG, gpu_modelcube, seed_offset=rank, out=pinned.numpy())
    cube_host = syn["cube"]                       # numpy view of the pinned buffer
    v = syn["v"]
    mm = MetaModel(fittype, 1)
    spec = SpecCube(cube_host, v)

    # ---------------- end-to-end arm: the public API from host buffers ----------------------------------------
    def e2e_step():
        uc = UltraCube(cube=spec, fittype=fittype, device=local)
        uc.fit_cube(ncomp=[1, 2], simpfit=True, guesses={1: syn["guess1"].copy(), 2: syn["guess2"].copy()})
        res = uc.get_best_2c_parcube()
        return uc, res

    # Steady state of a caller that processes one cube after another: the previous step's objects are released
    # before the next step starts.  (Keeping an older UltraCube alive costs ~7 ms per step in first-touch page
    # faults of the freshly mapped result arrays -- host allocator behaviour, not device work.)
    e2e_times = []
    n_e2e = max(2, min(args.steps, 5))
    for it in range(2 + n_e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        uc_i, res_i = e2e_step()
        torch.cuda.synchronize()
        if it >= 2:
            e2e_times.append(time.perf_counter() - t0)
        del uc_i, res_i
    e2e_t = torch.tensor([float(np.mean(e2e_times))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = npix * world / float(e2e_t.item())
    uc, res = e2e_step()                           # kept: the source of the rows and limits the device arm uses
    torch.cuda.synchronize()
    h2d = cube_host.nbytes + syn["guess1"].nbytes + syn["guess2"].nbytes + npix * 12 * 8
    d2h = sum(uc.pcubes[k].parcube.nbytes * 2 + npix * (8 + 8 + 8 + 8) for k in ("1", "2")) + npix * (3 + 1 + 3) * 8 \
        + npix

    # ---------------- device arm: rows resident in HBM --------------------------------------------------------
    rows = uc._ref_pcube.rows_on_device()
    p1h, p2h = uc.pcubes["1"], uc.pcubes["2"]
    # same guesses / limits the drivers produced (cubefit_simp clamps + velocity window)
    lims = {}
    for nc in (1, 2):
        g = syn["guess%d" % nc].copy()
        from mufasa_b200 import multi_v_fit as mvf
        vg = g[::4]
        vmed, v99, v1p = mvf.get_vstats(vg)
        vmax = max(vmed + 10, v99 + mm.sigmax)
        vmin = min(vmed - 10, v1p - mm.sigmax)
        lo = [vmin, mm.sigmin, mm.Texmin, mm.taumin] * nc
        hi = [vmax, mm.sigmax, mm.Texmax, mm.taumax] * nc
        g = synth.clamp_guesses(g, lo, hi, eps=mm.eps)
        lims[nc] = (lo, hi, torch.from_numpy(np.ascontiguousarray(g.reshape(4 * nc, npix).T)).to(dev))
    prob1 = _abi.make_problem(fittype, 1, nchan, npix, v[0], dv, lims[1][0], lims[1][1])
    prob2 = _abi.make_problem(fittype, 2, nchan, npix, v[0], dv, lims[2][0], lims[2][1])
    out1 = eng.fit(prob1, rows, lims[1][2])
    out2 = eng.fit(prob2, rows, lims[2][2])
    aout = eng.aicc_alloc(npix)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def dev_step(timers=None):
        if timers is not None:
            ev[0].record()
        # the 1- and the 2-component fit in flight together (b200fit_fit_batch): same results as two calls, the
        # latency-bound late rounds of one overlap the other
        eng.fit_batch([dict(prob=prob1, spectra=rows, guesses=lims[1][2], out=out1),
                       dict(prob=prob2, spectra=rows, guesses=lims[2][2], out=out2)])
        if timers is not None:
            ev[1].record()
            ev[2].record()
        eng.aicc_pass(prob1, rows, out1["params"], out2["params"], aout)
        best = eng.mask_argmax(aout["mask_counts"])
        bits = eng.mask_bits(prob1, out1["params"], out2["params"], best=best)
        eng.aicc_pass(prob1, rows, out1["params"], out2["params"], aout, fill_bits=bits, only_empty=True)
        if timers is not None:
            ev[3].record()

    for _ in range(max(3, args.warmup)):
        dev_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_fit1 = t_fit2 = t_aicc = 0.0
    launches0 = eng.launch_count()
    start.record()
    per = []
    for _ in range(args.steps):
        dev_step(timers=True)
        per.append(list(ev))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    stop.record()
    torch.cuda.synchronize()
    launches_timed = eng.launch_count() - launches0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = start.elapsed_time(stop)
    for e in per:
        t_fit1 += e[0].elapsed_time(e[1])
        t_fit2 += e[1].elapsed_time(e[2])
        t_aicc += e[2].elapsed_time(e[3])
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt.item()) / args.steps
    value = npix * world / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant kernel (fit_eval_kernel<2>) on rank 0 ---------------------------
    # One extra, untimed-for-`value` step with CUDA events recorded by the library around every kernel of the fit
    # (b200fit_set_profiling: events on the launching stream).
    peaks = measured_peaks()
    eng.set_profiling(True)
    eng.fit(prob1, rows, lims[1][2], out=out1)
    prof1 = eng.get_profile()
    eng.fit(prob2, rows, lims[2][2], out=out2)
    prof2 = eng.get_profile()
    eng.set_profiling(False)
    cnt = out2["counts"].to(torch.float64)
    evals = float(cnt[:, 0].sum().item())                 # fused model + Jacobian evaluations, summed over pixels
    exec_gauss = float(cnt[:, 2].sum().item()) * 32.0      # gaussians actually evaluated (windowed)
    exec_chan = float(cnt[:, 3].sum().item()) * 32.0       # channels actually visited (active 32-channel chunks)
    H, C, P = 18, 2, 8
    sfu = evals * nchan * C * (H + 2)                      # SFU_J = N*C*(H+2)
    f32 = evals * nchan * C * (12 * H + 30)                # F32_J = N*C*(12H+30)
    f64 = evals * nchan * (P * (P + 1) + 2 * P + 2)        # F64_J = N*(P(P+1)+2P+2)
    sfu_exec = exec_gauss + exec_chan * C                  # ex2 actually issued: gaussians + e^-tau per component
    launches_eval = max(1, prof2["rounds"])
    t_eval = prof2["eval_ms"] * 1e-3                       # sum over the launches of one fit
    clk = peaks["sm_max_mhz"] * 1e6
    nsm = eng.num_sms
    pk = {"fp32": nsm * 128 * 2 * clk, "xu": nsm * 16 * clk, "fp64": nsm * 64 * 2 * clk}
    ach = {"fp32": f32 / t_eval, "xu": sfu / t_eval, "fp64": f64 / t_eval}
    frac = {k: ach[k] / pk[k] for k in pk}
    bound = max(frac, key=lambda k: frac[k])
    hbm_bytes = npix * (nchan * 4 + P * 8 + (2 * P + 2) * 8 + 4 + 16)
    t2 = (prof2["prologue_ms"] + prof2["eval_ms"] + prof2["solve_ms"]) * 1e-3   # seconds per 2c fit run on its own
    roofline = {"bound": bound, "kernel": "fit_eval_kernel<2>", "achieved": ach[bound] / 1e12,
                "peak": pk[bound] / 1e12, "unit": "TFLOP/s" if bound != "xu" else "Top/s", "frac": frac[bound],
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch with all 65 536 pixels active (ncu --set
                # full, profiles/r1m_fit2c_ncu_full_summary.txt: 166.0 + 19.5 MB = 2831 B per pixel-evaluation: the
                # windowed part of the spectrum + the per-pixel LM state), scaled to the average launch of this run
                "traffic": NCU_DRAM_BYTES_PER_EVAL * evals / launches_eval,
                "traffic_source": "ncu capture r1m (2831 B per pixel-evaluation) x evaluations per launch of this run",
                "definition": "ALGORITHMIC work (SURVEY 8d per-evaluation figures x evaluations counted by the kernel) "
                              "/ summed CUDA-event duration of the kernel's launches in one 2c fit; the kernel skips "
                              "gaussians beyond 6.2 sigma (exactly 0 in fp32), see frac_executed",
                "frac_executed": {"xu": sfu_exec / t_eval / pk["xu"]},
                "pipes": {k: {"achieved_T": ach[k] / 1e12, "peak_T": pk[k] / 1e12, "frac": frac[k]} for k in pk},
                "peak_source": "derived: %d SMs x lanes x %.0f MHz (sm_max_mhz of MEASURED_PEAKS.json, %s)" % (
                    nsm, peaks["sm_max_mhz"], peaks["source"]),
                "hbm": {"achieved": hbm_bytes / t2 / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": hbm_bytes / t2 / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["source"]},
                "per_launch": {"launches": launches_eval, "avg_ms": prof2["eval_ms"] / launches_eval,
                               "xu_ops": sfu / launches_eval, "fp32_flop": f32 / launches_eval,
                               "fp64_flop": f64 / launches_eval},
                "algorithmic_per_fit": {"evaluations": evals, "xu_ops": sfu, "fp32_flop": f32, "fp64_flop": f64,
                                        "executed_gaussians": exec_gauss, "algorithmic_gaussians": evals * nchan * C * H},
                "fit_2c_ms": {"prologue": prof2["prologue_ms"], "eval": prof2["eval_ms"], "solve": prof2["solve_ms"],
                              "rounds": prof2["rounds"]},
                "fit_1c_ms": {"prologue": prof1["prologue_ms"], "eval": prof1["eval_ms"], "solve": prof1["solve_ms"],
                              "rounds": prof1["rounds"]},
                "share_of_step": {"fits_1c_and_2c_batched": t_fit1 / total_ms, "aicc": t_aicc / total_ms,
                                  "fit_eval_kernel<2>": prof2["eval_ms"] / ms_per_step,
                                  "fit_solve_kernel<2>": prof2["solve_ms"] / ms_per_step}}

    line = None
    if rank == 0:
        st2 = out2["status"].cpu().numpy()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 model+Jacobian, f64 normal equations/solve",
                "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": float(e2e_t.item()) * 1e3,
                        "samples_ms": [round(t * 1e3, 2) for t in e2e_times],
                        "api": "UltraCube.fit_cube([1, 2], simpfit) + get_best_2c_parcube()"},
                "gpu_launches": int(launches_timed),
                "fit2c_only": {"value": npix / t2, "unit": UNIT, "ms": t2 * 1e3},
                "roofline": roofline, "clocks": clocks,
                "fit_stats": {"converged_2c": float(np.isin(st2, [1, 2, 3, 4, 6]).mean()),
                              "maxiter_2c": float((st2 == 5).mean()),
                              "mean_evals_2c": evals / npix,
                              "mean_evals_1c": float(out1["counts"][:, 0].to(torch.float64).mean().item())}}
    if world > 1:
        dist.barrier()
    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rows_s = sample_rows(ny, 8)
        nspec, dt = cpu_job(CFG, rows_s, 8, cores)
        line["cpu_baseline"] = {"value": nspec / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d spectra (8 rows x every 8th column of the same cube), 1c fit + 2c fit + "
                                          "AICc with the numpy oracle (mpfit restatement), %d processes, %.1f s" % (
                                              nspec, cores, dt)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostics only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
