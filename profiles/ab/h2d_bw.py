"""Host -> device copy rates on this box: what bounds the upload of a cube slab (pinned, 1-D and 2-D strided)."""
import time
import torch
n = 1 << 28                     # 1 GiB of fp32
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
h.normal_()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("pinned 1-D 1 GiB", lambda: d.copy_(h, non_blocking=True)),
                 ("pinned 2-D [1000 x 128K of 256K]", lambda: d[:1000 * 131072].view(1000, 131072).copy_(
                     h[:1000 * 262144].view(1000, 262144)[:, :131072], non_blocking=True))):
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    nbytes = n * 4 if "1-D" in name else 1000 * 131072 * 4
    print("%s: %.1f GB/s" % (name, nbytes / dt / 1e9))
p = torch.empty(n // 4, dtype=torch.float32)
torch.cuda.synchronize(); t = time.perf_counter(); d[:n // 4].copy_(p); torch.cuda.synchronize()
print("pageable 256 MiB: %.1f GB/s" % (n / (time.perf_counter() - t) / 1e9))
torch.cuda.synchronize(); t = time.perf_counter(); h[:n // 4].copy_(d[:n // 4], non_blocking=True); torch.cuda.synchronize()
print("D2H pinned 256 MiB: %.1f GB/s" % (n / (time.perf_counter() - t) / 1e9))
