"""Separable convolution alone at configs[2] size: clean cube / cube with 20 blank rows, CUDA-event times."""
import os
import sys
sys.path.insert(0, os.getcwd())
import torch  # noqa: E402
from mufasa_b200 import convolve_tools as ct  # noqa: E402
from mufasa_b200.engine import get_engine  # noqa: E402

nchan, ny, nx = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1000, 512, 512)
reps = int(os.environ.get("REPS", "5"))
eng = get_engine(0)
g = torch.Generator(device="cuda").manual_seed(0)
cube = torch.randn((nchan, ny, nx), device="cuda", generator=g)
sep = ct.Beam(7.2, 7.2).deconvolve(ct.Beam(3.6, 3.6)).as_kernel(1.0)
kx, ky = (torch.from_numpy(sep[k]).cuda() for k in ("kx", "ky"))
dst = torch.empty_like(cube)
for name in ("clean", "blank20"):
    if name == "blank20":
        cube[:, :20] = float("nan")
    for mode in ("fast", "generic"):
        if mode == "generic":
            os.environ["B200FIT_CONV_GENERIC"] = "1"
        else:
            os.environ.pop("B200FIT_CONV_GENERIC", None)
        eng.convolve_planes(cube, kx=kx, ky=ky, out=dst)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            eng.convolve_planes(cube, kx=kx, ky=ky, out=dst)
            ev[i + 1].record()
        torch.cuda.synchronize()
        t = min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        gb = 2 * cube.numel() * 4 / 1e9
        print("%s %s K=%d: %.3f ms  %.0f GB/s" % (name, mode, kx.numel(), t, gb / t * 1e3))
