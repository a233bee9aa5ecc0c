"""Timing of the convolved-cube stage at the size of configs[2] (512 x 512 x 1000, GAS-like 3.6-pixel beam -> 23-pixel
kernel): CUDA events around b200fit_convolve_planes (separable and general path) and b200fit_resample_bilinear,
algorithmic HBM bytes (one read + one write of the cube) over the measured time, the oracle on a 4-channel sample
beside it.  Usage: python profiles/convolve_cfg3.py [nchan ny nx]"""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mufasa_b200 import convolve_tools as ct  # noqa: E402
from mufasa_b200.engine import get_engine  # noqa: E402
from oracle import convolve as oc  # noqa: E402

nchan, ny, nx = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1000, 512, 512)
eng = get_engine(0)
g = torch.Generator(device="cuda").manual_seed(0)
cube = torch.randn((nchan, ny, nx), device="cuda", generator=g)
cube[:, :20] = float("nan")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
hbm = float(peaks.get("hbm_gbs", 6554.9))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


out = {"shape": [nchan, ny, nx]}
sep = ct.Beam(7.2, 7.2).deconvolve(ct.Beam(3.6, 3.6)).as_kernel(1.0)
kx, ky = (torch.from_numpy(sep[k]).cuda() for k in ("kx", "ky"))
dst = torch.empty_like(cube)
t = timed(lambda: eng.convolve_planes(cube, kx=kx, ky=ky, out=dst))
gb = 2 * cube.numel() * 4 / 1e9
out["separable"] = dict(ksize=int(kx.numel()), ms=t, GBps=gb / t * 1e3, frac_hbm=gb / t * 1e3 / hbm)
ell = ct.Beam(7.2, 5.0, 30.0).deconvolve(ct.Beam(3.6, 2.5, 30.0)).as_kernel(1.0)
k2 = torch.from_numpy(ell["k2"]).cuda()
t = timed(lambda: eng.convolve_planes(cube, k2=k2, out=dst), reps=3)
K = int(k2.shape[0])
out["general"] = dict(ksize=K, ms=t, GBps=gb / t * 1e3, TFLOPs_fp32=4.0 * K * K * cube.numel() / t / 1e9)
t = timed(lambda: eng.resample_bilinear(dst, ny // 2, nx // 2, 2.0, 0.5, 2.0, 0.5))
gbr = (cube.numel() + cube.numel() // 4) * 4 / 1e9
out["resample"] = dict(ms=t, GBps=gbr / t * 1e3, frac_hbm=gbr / t * 1e3 / hbm)
# oracle (float64 numpy, one core) on a 4-channel sample
sample = cube[:4].cpu().numpy()
k64 = oc.beam_kernel(3.6, 3.6, 0.0, 1.0, 2)
t0 = time.perf_counter()
ref = oc.convolve_planes(sample, k64)
tc = time.perf_counter() - t0
out["cpu_oracle"] = dict(sample="4 channels", s=tc, Mvoxel_per_s=sample.size / tc / 1e6)
out["gpu_Mvoxel_per_s_separable"] = cube.numel() / out["separable"]["ms"] / 1e3
got = dst[:4]  # last convolution was the general kernel; redo the separable one for the check
eng.convolve_planes(cube, kx=kx, ky=ky, out=dst)
got = dst[:4].cpu().numpy()
ok = np.isfinite(ref)
out["max_abs_err_vs_oracle"] = float(np.max(np.abs(got[ok] - ref[ok])))
out["hbm_peak_GBps"] = hbm
print(json.dumps(out, indent=1))
