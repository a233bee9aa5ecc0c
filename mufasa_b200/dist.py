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


ROW_BLOCK = 8


def partition_row_blocks(ny, world, rank, block=ROW_BLOCK):
    """Rows of rank ``rank`` under the BLOCK-CYCLIC partition: blocks of ``block`` consecutive rows are dealt to the
    ranks round-robin (ascending row indices, int64 array).  Spectra differ a lot in cost (a 2-component fit of a
    faint or single-component spectrum takes several times the evaluations of a clean one) and the cost varies
    smoothly over the sky, so contiguous slabs leave the ranks unevenly loaded; interleaved blocks do not."""
    import numpy as np
    ny, world, rank, block = int(ny), int(world), int(rank), int(block)
    if world < 1 or not (0 <= rank < world) or block < 1:
        raise ValueError("bad world/rank/block")
    nblocks = -(-ny // block)
    rows = [np.arange(b * block, min(ny, (b + 1) * block)) for b in range(rank, nblocks, world)]
    return np.concatenate(rows).astype(np.int64) if rows else np.zeros(0, dtype=np.int64)


def owned_rows(ny, world, rank, partition="cyclic"):
    """Row indices of a rank: ``partition`` = 'cyclic' (partition_row_blocks) or 'slabs' (partition_rows)."""
    import numpy as np
    if partition == "slabs" or world == 1:
        y0, y1 = partition_rows(ny, world, rank)
        return np.arange(y0, y1, dtype=np.int64)
    if partition != "cyclic":
        raise ValueError("partition must be 'cyclic' or 'slabs'")
    return partition_row_blocks(ny, world, rank)


def take_rows(cube, rows):
    """[nchan, len(rows), nx] C-contiguous array of the given rows of ``cube``: a plain ndarray [nchan, ny, nx], or
    any object with a ``take_rows(rows)`` method (a loader that holds / reads only this rank's rows)."""
    import numpy as np
    if hasattr(cube, "take_rows"):
        return cube.take_rows(rows)
    if len(rows) and int(rows[-1]) - int(rows[0]) + 1 == len(rows):
        slab = cube[:, int(rows[0]):int(rows[-1]) + 1, :]
    else:
        slab = cube[:, rows, :]
    if not (isinstance(slab, np.ndarray) and slab.flags["C_CONTIGUOUS"]):
        slab = np.ascontiguousarray(slab)
    return slab


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


def _global_rank(group, r):
    """torch.distributed takes GLOBAL ranks for src / dst; ``r`` is a rank within ``group``."""
    return r if group is None else dist.get_global_rank(group, r)


def gather_maps(local, ny, group=None, dst=0):
    """Gather row-slab maps ``local[..., ry, nx]`` (any leading dims) to ``[..., ny, nx]`` on rank ``dst`` of the
    group (None elsewhere).  Slabs follow ``partition_rows``; uneven slabs are padded to the largest for the
    collective.  One collective whatever the number of maps: callers stack their maps along the leading dimension."""
    world, rank = _world(group)
    if world == 1:
        return local
    lead, nx = local.shape[:-2], local.shape[-1]
    rmax = -(-int(ny) // world)
    if local.shape[-2] == rmax:
        pad = local.contiguous()
    else:
        pad = torch.zeros(lead + (rmax, nx), dtype=local.dtype, device=local.device)
        pad[..., :local.shape[-2], :] = local
    if rank == dst:
        buf = torch.empty((world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
        pieces = list(buf.unbind(0))
    else:
        buf, pieces = None, None
    dist.gather(pad, pieces, dst=_global_rank(group, dst), group=group)
    if rank != dst:
        return None
    if int(ny) == world * rmax:          # even slabs: the gathered buffer is the result, rank-major along the rows
        nl = len(lead)
        perm = tuple(range(1, nl + 1)) + (0, nl + 1, nl + 2)
        return buf.permute(perm).reshape(lead + (int(ny), nx))
    out = torch.empty(lead + (int(ny), nx), dtype=local.dtype, device=local.device)
    for r, t in enumerate(pieces):
        y0, y1 = partition_rows(ny, world, r)
        out[..., y0:y1, :] = t[..., :y1 - y0, :]
    return out


def gather_rows(local, ny, partition="cyclic", group=None, dst=0):
    """``gather_maps`` for any row partition: ``local[..., ry, nx]`` holds the rows ``owned_rows(ny, world, rank,
    partition)`` of this rank; returns ``[..., ny, nx]`` on rank ``dst`` (None elsewhere).  One collective."""
    world, rank = _world(group)
    if world == 1 or partition == "slabs":
        return gather_maps(local, ny, group=group, dst=dst)
    lead, nx = local.shape[:-2], local.shape[-1]
    rmax = max(len(owned_rows(ny, world, r, partition)) for r in range(world))
    if local.shape[-2] == rmax:
        pad = local.contiguous()
    else:
        pad = torch.zeros(lead + (rmax, nx), dtype=local.dtype, device=local.device)
        pad[..., :local.shape[-2], :] = local
    if rank == dst:
        buf = torch.empty((world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
        pieces = list(buf.unbind(0))
    else:
        pieces = None
    dist.gather(pad, pieces, dst=_global_rank(group, dst), group=group)
    if rank != dst:
        return None
    out = torch.empty(lead + (int(ny), nx), dtype=local.dtype, device=local.device)
    for r, t in enumerate(pieces):
        rows = torch.from_numpy(owned_rows(ny, world, r, partition)).to(local.device)
        out.index_copy_(len(lead), rows, t[..., :rows.numel(), :])
    return out


def aicc_distributed(engine, prob, rows, params1, params2, first_pixel=0, expand=20, model_f32=True,
                     thr=(5.0, 5.0, 5.0), row_index=None, group=None, local_to_global=None):
    """``Engine.aicc_global`` across ranks: pass 1 on the local pixels, global arg-max of the un-dilated mask size,
    the owner evaluates that pixel's mask bits and broadcasts them, pass 2 redoes the local empty-mask pixels.
    ``first_pixel`` = global flat index of this rank's first pixel (y0 * nx) for a contiguous slab;
    ``local_to_global`` (callable: local flat index -> global flat index, monotonic) for any other partition."""
    world, rank = _world(group)
    out = engine.aicc_alloc(int(prob.npix))
    engine.aicc_pass(prob, rows, params1, params2, out, expand=expand, model_f32=model_f32, thr=thr,
                     row_index=row_index)
    if local_to_global is None:
        best = engine.mask_argmax(out["mask_counts"], index_offset=first_pixel)
        to_local = lambda g: g - first_pixel
    else:
        # the first local pixel holding the local maximum is also the one with the lowest GLOBAL index (the map is
        # monotonic); its global index enters the comparison across ranks
        best = engine.mask_argmax(out["mask_counts"], index_offset=0)
        c_loc, i_loc = (int(x) for x in best.tolist())
        best = torch.tensor([c_loc, int(local_to_global(i_loc)) if c_loc > 0 else 0], dtype=torch.int64,
                            device=best.device)
        to_local = lambda g, _i=i_loc: _i
    count, index, owner = global_best(best, group=group)
    if rank == owner and count > 0:
        bits = engine.mask_bits(prob, params1, params2, pix=to_local(index), model_f32=model_f32)
    else:
        bits = torch.zeros(64, dtype=torch.int32, device=engine.device)
    if world > 1:
        dist.broadcast(bits, src=_global_rank(group, owner), group=group)
    engine.aicc_pass(prob, rows, params1, params2, out, fill_bits=bits, only_empty=True, expand=expand,
                     model_f32=model_f32, thr=thr, row_index=row_index)
    out["fill_bits"] = bits
    out["fill_from"] = (count, index, owner)
    return out


def percentiles_like_numpy(values, qs):
    """``np.percentile(values, qs)`` (method 'linear') of a 1-D float64 torch tensor on ANY device, bit for bit:
    sort, virtual index (n - 1) q / 100, numpy's two-sided lerp.  Used for the global statistics of the velocity
    guesses (multi_v_fit.get_vstats, multi_v_fit.py:634-644), which every rank of a partitioned run needs and which
    cost 0.1 s per map of a 1024 x 1024 cube with numpy on the host."""
    import numpy as np
    n = int(values.numel())
    if n == 0:
        return [float("nan")] * len(qs)
    srt = torch.sort(values).values
    want, meta = [], []
    for q in qs:
        virt = (n - 1) * float(np.true_divide(q, np.float64(100)))
        prev = int(np.floor(virt))
        prev = min(max(prev, 0), n - 1)
        nxt = min(prev + 1, n - 1)
        want += [prev, nxt]
        meta.append(virt - np.floor(virt))
    got = srt[torch.tensor(want, device=srt.device)].cpu().numpy()
    out = []
    for k, t in enumerate(meta):
        a, b = np.float64(got[2 * k]), np.float64(got[2 * k + 1])
        d = b - a
        out.append(float(b - d * (1 - t)) if t >= 0.5 else float(a + d * t))
    return out


def median_like_numpy(values):
    """``np.median`` of a 1-D float64 torch tensor: the middle value, or the mean of the two middle values."""
    n = int(values.numel())
    if n == 0:
        return float("nan")
    k = torch.tensor([(n - 1) // 2, n // 2], device=values.device)
    lo, hi = torch.sort(values).values[k].tolist()
    return lo if n % 2 else (lo + hi) / 2.0


def copy_rows(dst, src, rows):
    """dst[:, i] = src[:, rows[i]] for [k, n, nx] arrays, by runs of consecutive rows (slices: plain memcpy; np.take
    along the middle axis is several times slower, and a cyclic partition owns its rows in blocks)."""
    import numpy as np
    rows = np.asarray(rows)
    a = 0
    while a < rows.size:
        b = a + 1
        while b < rows.size and rows[b] == rows[b - 1] + 1:
            b += 1
        dst[:, a:b] = src[:, int(rows[a]):int(rows[a]) + (b - a)]
        a = b


def vstats_of_device(t):
    """(median, 99th, 1st percentile) of the finite, non-zero entries of a float64 device tensor."""
    t = t.ravel()
    t = t[torch.isfinite(t) & (t != 0)]
    p99, p1 = percentiles_like_numpy(t, [99, 1])
    return median_like_numpy(t), p99, p1


def global_vstats(vmaps, device):
    """(median, 99th, 1st percentile) of the finite, non-zero velocity guesses ``vmaps`` [k, ny, nx] -- the numbers
    cubefit_simp derives its velocity window from (multi_v_fit.py:163-176) -- computed on ``device``."""
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(vmaps, dtype=np.float64)).to(device, non_blocking=True).ravel()
    t = t[torch.isfinite(t) & (t != 0)]
    p99, p1 = percentiles_like_numpy(t, [99, 1])
    return median_like_numpy(t), p99, p1


FIELDS = (("parcube1", 0, 4), ("errcube1", 4, 8), ("parcube2", 8, 16), ("errcube2", 16, 24), ("best_parcube", 24, 32),
          ("best_errcube", 32, 40), ("lnk10", 40, 41), ("lnk20", 41, 42), ("lnk21", 42, 43), ("ncomp", 43, 44))


class SelectionJob(object):
    """One rank's share of the model-selection job (see ``fit_and_select``), split at the points a caller may want
    to time or repeat: ``__init__`` = host work on the guesses + upload of the slab and the guesses; ``run`` =
    everything on the device, from the two fits to the gathered result block; ``to_host`` = the one copy back."""

    def __init__(self, cube, v_kms, fittype, guesses1, guesses2, device=None, group=None, vstats=None,
                 partition="cyclic"):
        import numpy as np
        from . import _abi
        from . import multi_v_fit as mvf
        from .cube import SpecCube
        from .engine import get_engine
        from .pcube import PCube
        from .spec_models.meta_model import MetaModel
        self.group = group
        self.world, self.rank = _world(group)
        nchan, ny, nx = cube.shape
        self.ny, self.nx, self.fittype = int(ny), int(nx), fittype
        self.partition = partition
        self.rows = rows = owned_rows(ny, self.world, self.rank, partition)
        self.ry = int(rows.size)
        self.npix = npix = self.ry * nx
        contiguous = self.ry == 0 or int(rows[-1]) - int(rows[0]) + 1 == self.ry
        dev = torch.cuda.current_device() if device is None else device
        self.eng = eng = get_engine(dev)
        slab = take_rows(cube, rows)
        spec = SpecCube(slab, v_kms)
        self.model_f32 = spec._data.dtype == np.float32
        self.ref = ref = PCube(spec._data, spec.xarr(), header=spec.header, unit=spec.unit, device=dev)
        # Host-to-device copies are served in the order they are queued, whatever their stream, so they are queued in
        # the order the first fit needs them: the velocity maps of the global statistics (24 MB), the FIRST slab of
        # the cube, this rank's guesses (staged on the host while that slab is crossing PCIe), then the other slabs.
        # Only then does the host wait for anything (the statistics: a sort on the GPU), and the kernels that need them
        # follow.  (Queued behind the whole cube, statistics and guesses arrived after it and nothing overlapped the
        # upload: 474 -> 402 ms end to end on one GPU at 1024 x 1024 x 1000, profiles/ab/e2e_cfg5_timeline.py.)
        main = torch.cuda.current_stream(eng.device)
        up = eng.side_stream("h2d")
        vdev = {}
        if vstats is None:
            for nc, full in ((1, guesses1), (2, guesses2)):
                vstage = eng.pinned_stage(("vmaps", nc, ny, nx))
                np.copyto(vstage.numpy(), full[::4])
                with torch.cuda.stream(up):
                    vdev[nc] = vstage.to(eng.device, non_blocking=True)
                    vdev[nc].record_stream(main)
            v_ready = torch.cuda.Event()
            v_ready.record(up)
        if npix:
            ref.rows_on_device(wait=False, max_slabs=1)
        # The front half of cubefit_simp (multi_v_fit.py:126-190) for both models, ON THE DEVICE: the rank's rows of the
        # guess maps go up as they are (through one pinned staging buffer, on a second stream, while the cube is
        # uploading), b200fit_prepare_guesses turns zero velocities into NaN and moves every value inside the limits,
        # and every pixel of the slab enters the fit -- the fit kernel itself reports a pixel with a non-finite guess
        # as "not fitted" (zeros), which is what the reference's maskmap & all-finite rule comes to.  On the host this
        # conditioning costs ~0.2 s per 1024 x 1024 map (copies, masks, clamps) and sat in front of the fits.
        raw = {}
        if npix:
            for nc, full in ((1, guesses1), (2, guesses2)):
                P = 4 * nc
                stage = eng.pinned_stage(("rawguess", P, self.ry, nx))
                if contiguous:
                    np.copyto(stage.numpy(), full[:, int(rows[0]):int(rows[-1]) + 1, :])
                else:
                    copy_rows(stage.numpy(), full, rows)
                with torch.cuda.stream(up):
                    raw[nc] = stage.to(eng.device, non_blocking=True).reshape(P, npix)
                    raw[nc].record_stream(main)
            ref.rows_on_device(wait=False)                   # the rest of the cube follows the guesses
        if vstats is None:
            main.wait_event(v_ready)
            vstats = {nc: vstats_of_device(vdev[nc]) for nc in (1, 2)}
        self.vstats = vstats
        self.batch, self.jobs = [], {}
        for nc in (1, 2):
            P = 4 * nc
            mm = MetaModel(fittype, nc)
            v_median, v_99p, v_1p = vstats[nc]
            if not npix or not np.isfinite(v_median):        # no finite velocity guess anywhere: nothing to fit
                self.jobs[nc] = None
                continue
            vmax, vmin = v_median + 10, v_median - 10        # multi_v_fit.py:163-176
            if v_99p + mm.sigmax > vmax:
                vmax = v_99p + mm.sigmax
            if v_1p - mm.sigmax < vmin:
                vmin = v_1p - mm.sigmax
            lo, hi = mm.limits(vmin, vmax)
            with torch.cuda.stream(up):
                g_dev = eng.prepare_guesses(raw[nc], lo, hi, mm.eps)
                ready = torch.cuda.Event()
                ready.record(up)
            g_dev.record_stream(main)
            prob = ref._problem(fittype, nc, npix, lo, hi, signal_cut=0.0)
            self.jobs[nc] = len(self.batch)
            self.batch.append(dict(prob=prob, spectra=ref._dev["rows"], guesses=g_dev, row_index=None,
                                   pcube=None, pix=None, ready=ready, slab_ready=ref._dev.get("slab_ready"),
                                   lo=lo, hi=hi))
        if npix:
            self.prob = ref._problem(fittype, 1, npix, [-1e30] * 4, [1e30] * 4)
        else:                                                # an empty slab still answers the collectives
            v0, dv, _ = ref.xarr.v0_dv_flip
            self.prob = _abi.make_problem(fittype, 1, nchan, 0, v0, dv, [-1e30] * 4, [1e30] * 4)
            self._no_rows = torch.zeros((0, eng.nchan_stride(nchan)), dtype=torch.float32, device=eng.device)

    def h2d_bytes(self):
        """Bytes this rank's ``__init__`` sends to its GPU: the slab and the guesses of the fitted pixels."""
        n = int(self.ref.cube.nbytes) if self.npix else 0
        for job in self.batch:
            n += int(job["guesses"].numel()) * 8 + (0 if job["row_index"] is None else int(job["row_index"].numel()) * 8)
        return n

    def _fit(self):
        """Both fits in flight together; results as {ncomp: dict(params, perr, ...)} in the slab's pixel order."""
        eng, npix = self.eng, self.npix
        outs = eng.fit_batch(self.batch) if self.batch else []
        fits = {}
        for nc in (1, 2):
            P = 4 * nc
            if self.jobs[nc] is None:
                fits[nc] = dict(params=torch.zeros((npix, P), dtype=torch.float64, device=eng.device),
                                perr=torch.zeros((npix, P), dtype=torch.float64, device=eng.device))
                continue
            out, job = outs[self.jobs[nc]], self.batch[self.jobs[nc]]
            if job["row_index"] is None:
                fits[nc] = out
            else:                                            # masked fit: scatter to the slab's pixel order
                full_p = torch.zeros((npix, P), dtype=torch.float64, device=eng.device)
                full_e = torch.zeros((npix, P), dtype=torch.float64, device=eng.device)
                full_p[job["row_index"]] = out["params"]
                full_e[job["row_index"]] = out["perr"]
                fits[nc] = dict(params=full_p, perr=full_e, status=out["status"], counts=out.get("counts"))
        self.fits = fits
        return fits

    def _select_and_pack(self, fits):
        """The AICc selection (global include_nosamp rule) and the assembly of this rank's [44, rows, nx] block."""
        from . import _abi
        eng, npix = self.eng, self.npix
        rows = self.ref.rows_on_device() if npix else self._no_rows
        nx, owned = self.nx, self.rows
        self.sel = sel = aicc_distributed(eng, self.prob, rows, fits[1]["params"], fits[2]["params"],
                                          model_f32=self.model_f32, group=self.group,
                                          local_to_global=lambda i: int(owned[i // nx]) * nx + i % nx)
        return eng.pack_selection(fits[1], fits[2], sel).reshape(_abi.PACK_FIELDS, self.ry, self.nx)

    def _gather(self, block):
        if dist.get_backend(self.group) != "nccl":
            block = block.cpu()
        return gather_rows(block, self.ny, partition=self.partition, group=self.group, dst=0)

    def run(self, gather=True):
        """Both fits in flight together, the AICc selection and the assembly of the result block on the fit results
        where they lie, then ONE gather.  Returns the [44, ny, nx] block on rank 0 (None elsewhere), or this rank's
        [44, rows, nx] block when ``gather`` is False.  Device tensors; no copy to the host except the 16 bytes per
        rank of the global arg-max."""
        packed = self._select_and_pack(self._fit())
        if gather and self.world > 1:
            packed = self._gather(packed)
        return packed

    def to_host(self, packed):
        """The result block as a dict of numpy views of ONE pinned host buffer (valid until the next call)."""
        if packed is None:
            return None
        stage = self.eng.pinned_stage(("select",) + tuple(packed.shape))
        stage.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(self.eng.device).synchronize()
        host = stage.numpy()
        return {name: (host[a] if b - a == 1 else host[a:b]) for name, a, b in FIELDS}

    def run_to_host(self, gather=True):
        """``to_host(run())`` with the read-back split in two: the fit maps (24 of the 44) are final when the fits are,
        so their gather and their copy to the host run on a second stream UNDER the selection pass; only the 20
        selection maps follow it.  Same values, same pinned result buffer."""
        eng = self.eng
        main = torch.cuda.current_stream(eng.device)
        side = eng.side_stream("d2h")
        fits = self._fit()
        nfit = FIELDS[3][2]                                    # parcube1, errcube1, parcube2, errcube2
        collect = gather and self.world > 1
        out_rows = self.ny if collect else self.ry
        stage = eng.pinned_stage(("select", FIELDS[-1][2], out_rows, self.nx))
        fits_done = torch.cuda.Event()
        fits_done.record(main)
        with torch.cuda.stream(side):
            side.wait_event(fits_done)
            part = torch.cat([fits[1]["params"].t(), fits[1]["perr"].t(), fits[2]["params"].t(), fits[2]["perr"].t()]
                             ).reshape(nfit, self.ry, self.nx)
            if collect:
                part = self._gather(part)
            if part is not None:
                part.record_stream(side)
                stage[:nfit].copy_(part, non_blocking=True)
        packed = self._select_and_pack(fits)[nfit:]
        if collect:
            packed = self._gather(packed)
        if packed is not None:
            stage[nfit:].copy_(packed, non_blocking=True)
        side.synchronize()
        main.synchronize()
        if packed is None:
            return None
        host = stage.numpy()
        return {name: (host[a] if b - a == 1 else host[a:b]) for name, a, b in FIELDS}


def fit_and_select(cube, v_kms, fittype, guesses1, guesses2, device=None, group=None, gather=True, vstats=None,
                   partition="cyclic"):
    """The model-selection job (1c fit + 2c fit + AICc 0/1/2 selection; UltraCube.fit_cube([1, 2], simpfit=True) +
    get_best_2c_parcube) on a cube partitioned by blocks of rows over the ranks of ``group``: ``partition`` =
    'cyclic' (blocks of ROW_BLOCK rows dealt round-robin: even load, see partition_row_blocks) or 'slabs' (one
    contiguous slab per rank).

    ``cube`` [nchan, ny, nx] (or a loader object with ``shape`` and ``take_rows(rows)``), ``guesses1`` [4, ny, nx],
    ``guesses2`` [8, ny, nx]: the WHOLE arrays on every rank (each rank touches only its rows).
    Global couplings are resolved before partitioning: cubefit_simp's velocity window comes from the statistics of
    the whole guess map (``vstats`` = {1: (median, p99, p1), 2: ...} if the caller has them, else computed here on
    the GPU), the include_nosamp fill mask from ``aicc_distributed``.  Per-pixel results are identical to a
    single-GPU run.

    Device-resident from the upload to the gather: the slab goes up (pinned host memory: asynchronously, in two
    pieces, the first being fitted while the second crosses PCIe), both fits run in flight together, the AICc
    selection and the assembly of the best-model cube (b200fit_pack_selection) read the fit results where they lie,
    ONE NCCL gather brings the [44, rows, nx] result block of every rank to rank 0, ONE copy brings it to the host.
    A slab that cannot be fitted at all (no finite guess: SNRMaskError; no pixel yields a fit) comes back as
    "nothing fitted" -- the rank still takes part in every collective.

    Returns dict(parcube1, errcube1, parcube2, errcube2, best_parcube, best_errcube, lnk10, lnk20, lnk21, ncomp) of
    [.., ny, nx] numpy arrays on rank 0 (None elsewhere) when ``gather``, else the slab's.  The arrays are views of
    a pinned staging buffer that the next call on this device reuses: copy what must outlive it."""
    import os
    import time
    stamps = [time.perf_counter()] if os.environ.get("B200FIT_TIMELINE") else None
    job = SelectionJob(cube, v_kms, fittype, guesses1, guesses2, device=device, group=group, vstats=vstats,
                       partition=partition)
    if stamps is not None:
        stamps.append(time.perf_counter())
    res = job.run_to_host(gather=gather)
    if stamps is not None:                     # host work + queued uploads | fits, selection, gather and read-back
        stamps.append(time.perf_counter())
        import sys
        sys.stderr.write("fit_and_select rank %d: init %.1f run_to_host %.1f ms\n" % (
            (job.rank,) + tuple(1e3 * (b - a) for a, b in zip(stamps[:-1], stamps[1:]))))
    return res
