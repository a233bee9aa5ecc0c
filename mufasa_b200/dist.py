"""Multi-GPU composition of the fit path: one process per GPU, pixels partitioned by row slabs.

Pixels are independent once guesses, limits and masks are fixed, so there is NO collective on the
data path: every rank fits its own slab with the single-GPU calls.  What crosses ranks:
  * the result maps, gathered once per fit call (``gather_maps``; NCCL on GPUs, gloo in the CPU tests);
  * the ``include_nosamp`` rule of the AICc pass (UltraCube.py:1423-1438): pixels whose own model mask is empty
    borrow the spectral mask of the FIRST (row-major) pixel holding the largest mask -- a global arg-max
    (``global_best``) followed by a broadcast of that pixel's 64-word bit set (``aicc_distributed``).
Host-side global scalars (velocity limits multi_v_fit.py:253-254, the m0 percentile moment_guess.py:461-467, guess
medians multi_v_fit.py:634-644) must be computed over the whole cube BEFORE partitioning; the drivers in
multi_v_fit.py take them as arguments for that reason.

torch.distributed is plumbing only; nothing here computes on the fit path.
"""
import torch
import torch.distributed as dist


def partition_rows(ny, world, rank):
    """Contiguous row slab [y0, y1) of rank ``rank``: the first ny % world ranks get one extra row."""
    ny, world, rank = int(ny), int(world), int(rank)
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(ny, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def global_best(local_best, group=None):
    """``local_best`` = int64[2] tensor {largest mask count on this rank, GLOBAL flat index of the first pixel
    holding it} (b200fit_mask_argmax with index_offset = first pixel of the slab).  Returns (count, index, owner)
    of the global winner: the largest count, ties broken by the lowest global index -- what
    ``np.where(nsamp == nanmax(nsamp))`` picks first on the undivided cube (UltraCube.py:1428-1429)."""
    world, rank = _world(group)
    if world == 1:
        c, i = (int(x) for x in local_best.tolist())
        return c, i, 0
    pieces = [torch.empty_like(local_best) for _ in range(world)]
    dist.all_gather(pieces, local_best.contiguous(), group=group)
    best = (-1, -1, -1)
    for r, t in enumerate(pieces):
        c, i = (int(x) for x in t.tolist())
        if c > best[0] or (c == best[0] and c > 0 and i < best[1]):
            best = (c, i, r)
    return best


def gather_maps(local, ny, group=None, dst=0):
    """Gather row-slab maps ``local[..., ry, nx]`` (any leading dims) to ``[..., ny, nx]`` on rank ``dst``
    (None elsewhere).  Slabs follow ``partition_rows``; uneven slabs are padded to the largest for the collective."""
    world, rank = _world(group)
    if world == 1:
        return local
    lead, nx = local.shape[:-2], local.shape[-1]
    rmax = -(-int(ny) // world)
    pad = torch.zeros(lead + (rmax, nx), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-2], :] = local
    pieces = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, pieces, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty(lead + (int(ny), nx), dtype=local.dtype, device=local.device)
    for r, t in enumerate(pieces):
        y0, y1 = partition_rows(ny, world, r)
        out[..., y0:y1, :] = t[..., :y1 - y0, :]
    return out


def aicc_distributed(engine, prob, rows, params1, params2, first_pixel, expand=20, model_f32=True,
                     thr=(5.0, 5.0, 5.0), row_index=None, group=None):
    """``Engine.aicc_global`` across ranks: pass 1 on the local pixels, global arg-max of the un-dilated mask size,
    the owner evaluates that pixel's mask bits and broadcasts them, pass 2 redoes the local empty-mask pixels.
    ``first_pixel`` = global flat index of this rank's first pixel (y0 * nx)."""
    world, rank = _world(group)
    out = engine.aicc_alloc(int(prob.npix))
    engine.aicc_pass(prob, rows, params1, params2, out, expand=expand, model_f32=model_f32, thr=thr,
                     row_index=row_index)
    best = engine.mask_argmax(out["mask_counts"], index_offset=first_pixel)
    count, index, owner = global_best(best, group=group)
    if rank == owner and count > 0:
        bits = engine.mask_bits(prob, params1, params2, pix=index - first_pixel, model_f32=model_f32)
    else:
        bits = torch.zeros(64, dtype=torch.int32, device=engine.device)
    if world > 1:
        dist.broadcast(bits, src=owner, group=group)
    engine.aicc_pass(prob, rows, params1, params2, out, fill_bits=bits, only_empty=True, expand=expand,
                     model_f32=model_f32, thr=thr, row_index=row_index)
    out["fill_bits"] = bits
    out["fill_from"] = (count, index, owner)
    return out


def fit_and_select(cube, v_kms, fittype, guesses1, guesses2, device=None, group=None, gather=True):
    """The model-selection job (1c fit + 2c fit + AICc 0/1/2 selection; UltraCube.fit_cube([1, 2]) +
    get_best_2c_parcube) on a cube partitioned by row slabs over the ranks of ``group``.

    ``cube`` [nchan, ny, nx], ``guesses1`` [4, ny, nx], ``guesses2`` [8, ny, nx]: the WHOLE arrays on every rank
    (each rank touches only its slab; a loader that reads slabs directly can pass views of a larger virtual array).
    Global couplings are resolved before partitioning: cubefit_simp's velocity window comes from the statistics of
    the whole guess map, the include_nosamp fill mask from ``aicc_distributed``.  Per-pixel results are identical
    to a single-GPU run.  Returns dict(parcube1, errcube1, parcube2, errcube2, best_parcube, best_errcube, lnk10,
    lnk20, lnk21, ncomp) of [.., ny, nx] numpy arrays on rank 0 (None elsewhere) when ``gather``, else the slab's."""
    import numpy as np
    from . import multi_v_fit as mvf
    from .cube import SpecCube
    from .UltraCube import UltraCube
    world, rank = _world(group)
    nchan, ny, nx = cube.shape
    y0, y1 = partition_rows(ny, world, rank)
    dev = torch.cuda.current_device() if device is None else device
    slab = np.ascontiguousarray(cube[:, y0:y1, :])
    uc = UltraCube(cube=SpecCube(slab, v_kms), fittype=fittype, device=dev)
    uc.dist = (group, y0 * nx)
    g = {1: np.array(guesses1[:, y0:y1, :], dtype=np.float64), 2: np.array(guesses2[:, y0:y1, :], dtype=np.float64)}
    vstats = {}
    for nc, full in ((1, guesses1), (2, guesses2)):      # global statistics of the velocity guesses (cubefit_simp)
        vg = np.array(full[::4], dtype=np.float64)
        vg[vg == 0] = np.nan
        vstats[nc] = mvf.get_vstats(vg)
    for nc in (1, 2):
        uc.pcubes.pop(str(nc), None)
    # the two fits in flight together; each driver gets the global velocity statistics of its own guess map
    batch = []
    for nc in (1, 2):
        p = uc.load_pcube(pcube_ref=uc._ref_pcube)
        p._deferred = batch
        uc.pcubes[str(nc)] = mvf.cubefit_simp(uc.cube, p, ncomp=nc, guesses=g[nc], fittype=fittype,
                                              vstats=vstats[nc])
        p._deferred = None
    if batch:
        outs = batch[0]["pcube"].engine.fit_batch(batch)
        for job, out in zip(batch, outs):
            job["pcube"].finish_fit(job, out)
    for nc in (1, 2):
        uc._include_fit_in_mask(nc)
    best_par, best_err, lnk10, lnk20, lnk21 = uc.get_best_2c_parcube()
    local = dict(parcube1=uc.pcubes["1"].parcube, errcube1=uc.pcubes["1"].errcube, parcube2=uc.pcubes["2"].parcube,
                 errcube2=uc.pcubes["2"].errcube, best_parcube=best_par, best_errcube=best_err,
                 lnk10=lnk10[None], lnk20=lnk20[None], lnk21=lnk21[None],
                 ncomp=uc.ncomp_map.astype(np.float64)[None])
    if not gather or world == 1:
        return {k: (v[0] if k.startswith("lnk") or k == "ncomp" else v) for k, v in local.items()}
    out = {}
    device_t = torch.device("cuda", dev) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    for k, v in local.items():
        t = gather_maps(torch.from_numpy(np.ascontiguousarray(v)).to(device_t), ny, group=group, dst=0)
        if rank == 0:
            a = t.cpu().numpy()
            out[k] = a[0] if k.startswith("lnk") or k == "ncomp" else a
    return out if rank == 0 else None
