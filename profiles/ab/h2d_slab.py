"""Upload rate of a row slab of a pinned [nchan, plane] cube: b200fit_upload_slab (one 2-D copy per 16 MB piece) against
one 1-D copy per channel row, for the slab shapes of the bench (whole cube in 2 / 4 slabs; an eighth of it in 2)."""
import sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mufasa_b200.engine import get_engine
eng = get_engine(0)
nchan = 1000
for plane, nslab in ((1048576, 2), (1048576, 4), (131072, 2)):
    t = torch.empty((nchan, plane), dtype=torch.float32, pin_memory=True)
    t.fill_(1.0)
    count = plane // nslab
    staging = torch.empty((nchan, count), dtype=torch.float32, device=eng.device)
    def two_d():
        step = max(1, (16 << 20) // (4 * count))
        for c0 in range(0, nchan, step):
            c1 = min(nchan, c0 + step)
            eng.upload_slab(t[c0:c1], 0, count, staging[c0:c1])
    def one_d():
        for c in range(nchan):
            staging[c].copy_(t[c, :count], non_blocking=True)
    def whole_2d():
        eng.upload_slab(t, 0, count, staging)
    for name, fn in (("2-D, 16 MB pieces", two_d), ("2-D, one call", whole_2d), ("1-D per channel row", one_d)):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); th = time.perf_counter() - t0
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        print("plane %8d, slab %7d px (%5.0f MB): %-20s %6.1f GB/s (host time of the calls %.1f ms)" % (
            plane, count, nchan * count * 4 / 1e6, name, nchan * count * 4 / best / 1e9, th * 1e3), flush=True)
    del t, staging
