"""Ensemble sharding across the GPUs of one box (one process per GPU).

Reactor instances are independent initial-value problems — nothing couples them
(the reference solves exactly one per `rmtExe` call, PyREMOT/docs/rmtCore.py:393-413) —
so the ensemble is split into contiguous index blocks, one per rank, and there is NO
collective on the data path.  The cross-GPU step happens once, at the end of a run,
and is ONE collective: every rank lays its results out in a single FP64 buffer

    [ result rows (outlet rows | objectives | profiles) ][ status (int32) ][ pad ][ tail: sum, min, argmin, failed ]

(`PackLayout`) which the integrator kernel writes directly — outlets, objectives, status
and, through the fused last-block reduction of `rmt_n1_solve_population`, the tail — and
which is all-gathered over NVLink in one call.  No host synchronisation happens before
the collective; the host reads the gathered buffer once.

Transport: NCCL, either through `torch.distributed` (`group`) or through the C ABI's own
`rmt_comm_*` entry points (`comm`, a `capi.Comm`: no torch.distributed needed).  The
helpers only move tensors; they run unchanged on CPU tensors under gloo, which is how the
host logic is tested without a GPU (tests/test_distributed_cpu.py).
"""
import numpy as np


def partition(B, world, rank):
    """Contiguous block split of range(B): the first B % world ranks get one extra."""
    base, extra = divmod(int(B), int(world))
    start = rank*base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sweep(sweep, B, world, rank):
    lo, hi = partition(B, world, rank)
    return {k: v[lo:hi] for k, v in (sweep or {}).items()}, lo, hi


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None, comm=None):
    if comm is not None:
        return comm.rank, comm.nranks
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


# ----------------------------------------------------------------------------------
# the packed per-rank buffer
# ----------------------------------------------------------------------------------
class PackLayout:
    """Offsets (in doubles) of one rank's packed result buffer.  `nrows` FP64 rows of the rank's own length
    B_r come first (so a kernel that writes [rows][B_r] with row stride B_r can write straight into it), then
    the int32 status words, then padding up to the common length, and the 4-double tail at the very end (a
    fixed offset for every rank)."""
    TAIL = 4

    def __init__(self, B, world, nrows):
        self.B, self.world, self.nrows = int(B), int(world), int(nrows)
        self.sizes = [partition(B, world, r)[1] - partition(B, world, r)[0] for r in range(world)]
        self.starts = [partition(B, world, r)[0] for r in range(world)]
        bmax = max(self.sizes) if self.sizes else 0
        self.length = self.nrows*bmax + (bmax + 1)//2 + self.TAIL

    def views(self, buf, rank):
        """(rows [nrows, B_r] float64, status [B_r] int32, tail [4] float64) — views into this rank's buffer."""
        import torch
        Br = self.sizes[rank]
        rows = buf[:self.nrows*Br].view(self.nrows, Br)
        st = buf[self.nrows*Br:self.nrows*Br + (Br + 1)//2].view(torch.int32)[:Br]
        return rows, st, buf[self.length - self.TAIL:]

    def unpack_device(self, gathered, ws, transpose=True):
        """Device-side form of `unpack` for large gathers: the per-rank blocks are scattered (and transposed to
        instance-major [B, nrows] when `transpose`) by the GPU, and rows, status and tails travel to the host in
        ONE copy each into pinned workspace memory — instead of a pageable D2H of the packed buffer followed by two
        host passes over it (measured on 2^23 reactors: 0.63 s -> see bench strong_scaling).  Returns host numpy
        views (valid until the next call with the same workspace): rows [B, nrows] (or [nrows, B]), status [B],
        tails [world, 4]."""
        import torch
        dev = gathered.device
        g = gathered.view(self.world, self.length)
        shape = (self.B, self.nrows) if transpose else (self.nrows, self.B)
        d_rows = ws.get("d_unpack_rows", shape, torch.float64, device=dev)
        d_st = ws.get("d_unpack_status", (self.B,), torch.int32, device=dev)
        # (one 2-D transposing copy per rank: a single batched permute-copy over all ranks was measured 5x slower — torch's
        # strided-copy kernel writes 9-double rows uncoalesced — 11 ms instead of 2.2 ms per 65 536-set population)
        for r in range(self.world):
            Br, lo = self.sizes[r], self.starts[r]
            if Br == 0:
                continue
            blk = g[r, :self.nrows*Br].view(self.nrows, Br)
            if transpose:
                d_rows[lo:lo + Br].copy_(blk.t())
            else:
                d_rows[:, lo:lo + Br].copy_(blk)
            d_st[lo:lo + Br].copy_(g[r, self.nrows*Br:self.nrows*Br + (Br + 1)//2].view(torch.int32)[:Br])
        h_rows = ws.get("h_unpack_rows", shape, torch.float64, pinned=True)
        h_st = ws.get("h_unpack_status", (self.B,), torch.int32, pinned=True)
        h_tail = ws.get("h_unpack_tails", (self.world, self.TAIL), torch.float64, pinned=True)
        h_rows.copy_(d_rows, non_blocking=True)
        h_st.copy_(d_st, non_blocking=True)
        h_tail.copy_(g[:, self.length - self.TAIL:], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return h_rows.numpy(), h_st.numpy(), h_tail.numpy().copy()

    def unpack(self, gathered):
        """gathered [world, length] (host numpy float64) -> rows [nrows, B], status [B] int32, tails [world, 4]."""
        g = np.ascontiguousarray(gathered).reshape(self.world, self.length)
        rows = np.empty((self.nrows, self.B), dtype=np.float64)
        status = np.empty((self.B,), dtype=np.int32)
        for r in range(self.world):
            Br, lo = self.sizes[r], self.starts[r]
            if Br == 0:
                continue
            rows[:, lo:lo + Br] = g[r, :self.nrows*Br].reshape(self.nrows, Br)
            status[lo:lo + Br] = g[r, self.nrows*Br:self.nrows*Br + (Br + 1)//2].view(np.int32)[:Br]
        return rows, status, g[:, self.length - self.TAIL:].copy()


def fold_tails(tails):
    """(sum, min, argmin, failed) of the whole ensemble from the per-rank tails [world, 4], in rank order
    (deterministic); ties on the minimum resolve to the smallest global index."""
    t = np.asarray(tails, dtype=np.float64).reshape(-1, 4)
    total = float(np.sum(t[:, 0]))
    valid = t[:, 2] >= 0
    if not valid.any():
        return total, float("inf"), -1, int(round(t[:, 3].sum()))
    mn = t[valid, 1].min()
    cand = t[valid & (t[:, 1] == mn), 2]
    return total, float(mn), int(cand.min()), int(round(t[:, 3].sum()))


def all_gather_packed(buf, world, group=None, comm=None, stream=None):
    """ONE all-gather of every rank's packed buffer: [length] -> [world, length] on every rank."""
    import torch
    if world == 1:
        return buf.view(1, -1)
    out = torch.empty((world, buf.numel()), dtype=buf.dtype, device=buf.device)
    if comm is not None:
        comm.allgather(buf, out, buf.numel(), stream=stream)
    else:
        _dist().all_gather_into_tensor(out.view(-1), buf, group=group)
    return out


# ----------------------------------------------------------------------------------
# small helpers kept for callers that gather single arrays
# ----------------------------------------------------------------------------------
def reduce_objective(local_sum, local_min, local_argmin, device=None, group=None):
    """Global (sum, min, argmin) of a sharded objective from the per-rank triples (one all-gather of 3 doubles)."""
    import torch
    rank, world = world_info(group)
    if world == 1:
        return float(local_sum), float(local_min), int(local_argmin)
    t = torch.tensor([float(local_sum), float(local_min), float(local_argmin), 0.0], dtype=torch.float64, device=device)
    allt = all_gather_packed(t, world, group).cpu().numpy()
    return fold_tails(allt)[:3]


def all_gather_rows(local, B, group=None):
    """Concatenate per-rank shards along the LAST dimension (the instance index) into
    the full ensemble on every rank.  `local`: tensor [..., B_local]."""
    import torch
    rank, world = world_info(group)
    if world == 1:
        return local
    sizes = [partition(B, world, r)[1] - partition(B, world, r)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros(local.shape[:-1] + (mx,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    out = torch.empty((world*pad.numel(),), dtype=local.dtype, device=local.device)
    _dist().all_gather_into_tensor(out, pad.contiguous().view(-1), group=group)
    out = out.view((world,) + tuple(pad.shape))
    return torch.cat([out[r][..., :sizes[r]] for r in range(world)], dim=-1)


def _batch_size(sweep, B):
    if B is not None:
        return int(B)
    first = next(iter(sweep.values()))
    return int(first.shape[0]) if hasattr(first, "shape") else len(first)


# ----------------------------------------------------------------------------------
# steady-state models (N1, M7)
# ----------------------------------------------------------------------------------
def rmtExeBatchSharded(modelInput, sweep, B=None, *, rtol=None, atol=None, objective_ref=None, gather=True,
                       group=None, comm=None, workspace=None, zNo=None, tNo=None):
    """Every rank calls this with the SAME full `sweep`; each solves its block on its own GPU.

    Models N1 / M7 (outlet of every reactor); N2 / M9 are forwarded to `rmtExeBatchN2Sharded`.
    Returns a dict with the rank's shard ("local_*", device tensors — views into the packed buffer, valid until the
    next call with the same workspace), "failed" (reactors with status != 0 in the whole ensemble), the global
    objective statistics when `objective_ref` is given, and — with `gather=True` — the full outlet array [B, n],
    status [B] and objectives [B] (host numpy) on every rank (`gather="root"`: on rank 0 only; the other ranks take
    part in the collective and read back nothing but the 4-number tails).  One collective per call."""
    import torch
    from . import engine
    from .rmt import _check_components
    _check_components(modelInput)
    if modelInput["model"] in ("N2", "M9"):
        return rmtExeBatchN2Sharded(modelInput, sweep, B, zNo=zNo, tNo=tNo, rtol=rtol, atol=atol,
                                    gather="final" if gather else None, group=group, comm=comm, workspace=workspace)
    B = _batch_size(sweep, B)
    rank, world = world_info(group, comm)
    local, lo, hi = shard_sweep(sweep, B, world, rank)
    Bl = hi - lo
    sc = modelInput.get('solver-config', {})
    rtol = float(sc.get('rtol', engine.DEFAULT_RTOL) if rtol is None else rtol)
    atol = float(sc.get('atol', engine.DEFAULT_ATOL) if atol is None else atol)
    cm = engine.compile_model(modelInput, method=engine.choose_method(modelInput, rtol, 1))
    if not torch.cuda.is_available():
        from . import capi
        raise capi.RmtError("rmt_app_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    mod = cm.load(dev.index)
    spec, n = cm.spec, cm.spec.n
    ws = workspace if workspace is not None else engine.Workspace()
    with_obj = objective_ref is not None
    lay = PackLayout(B, world, n + (1 if with_obj else 0))
    ctrl = engine.METHOD_CTRL.get(cm.method)
    z_end = float(modelInput['reactor']['ReLe']) if spec.model == "M7" else 1.0
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        pack = ws.get("d_pack", (lay.length,), torch.float64, device=dev)
        rows, status, tail = lay.views(pack, rank)
        d_stats = ws.get("d_stats", (4, max(Bl, 1)), torch.int32, device=dev)
        if Bl > 0:
            d_rows, n_rows, row_map, _ = engine.sweep_rows_to_device(spec, local, Bl, ws, dev)
            d_consts = ws.get("d_consts", (mod.info.nconst, Bl), torch.float64, device=dev)
            mod.setup(Bl, d_rows, n_rows, row_map, engine.uniform_inputs(spec, modelInput), d_consts, stream=stream)
            if with_obj:
                mod.n1_solve_population(Bl, d_consts, z_end, rtol, atol, rows[:n], status, d_stats, objective_ref, rows[n],
                                        tail, index_offset=lo, ctrl=ctrl, stream=stream)
            else:
                mod.n1_solve(Bl, d_consts, np.array([z_end]), rtol, atol, rows[:n], status, d_stats, dense=False,
                             out_mode=1, ctrl=ctrl, stream=stream)
                tail.zero_()
                tail[2] = -1.0
                tail[3] = (status != 0).sum()
        else:
            tail.copy_(torch.tensor([0.0, float("inf"), -1.0, 0.0], dtype=torch.float64), non_blocking=True)
        out = {"range": (lo, hi), "local_dataYs": rows[:n], "local_status": status, "local_stats": d_stats[:, :Bl],
               "local_objective": rows[n] if with_obj else None}
        if gather == "root" and rank != 0:
            g = all_gather_packed(pack, world, group, comm, stream)                        # take part in the collective;
            tails = g[:, lay.length - lay.TAIL:].cpu().numpy()                             # only the tails go to this host
        elif gather:
            g = all_gather_packed(pack, world, group, comm, stream)
            if g.is_cuda:
                # scatter + transpose on the device, one D2H into pinned memory (views of the workspace)
                full, st, tails = lay.unpack_device(g, ws, transpose=True)                 # [B, n (+1)]
                out["dataYs"] = full[:, :n]
                out["status"] = st
                if with_obj:
                    out["objective"] = full[:, n]
            else:                                                                          # CPU tensors (gloo tests)
                full, st, tails = lay.unpack(g.numpy())
                out["dataYs"] = np.ascontiguousarray(full[:n].T)
                out["status"] = st
                if with_obj:
                    out["objective"] = full[n].copy()
        else:
            tails = all_gather_packed(tail, world, group, comm, stream).cpu().numpy()
        s, mn, am, bad = fold_tails(tails)
        out["failed"] = bad
        if with_obj:
            out["objective_sum"], out["objective_min"], out["objective_argmin"] = s, mn, am
    return out


# ----------------------------------------------------------------------------------
# dynamic models (N2, M9)
# ----------------------------------------------------------------------------------
def rmtExeBatchN2Sharded(modelInput, sweep, B=None, *, zNo=None, tNo=None, rtol=None, atol=None, gather="final",
                         group=None, comm=None, workspace=None, keep_on_device=False):
    """Sharded form of `rmtExeBatchN2` (BASELINE config 5: 200 nodes x 100 000 reactors over 8 GPUs): every rank
    integrates its contiguous block of reactors; axial nodes are never split across GPUs.

    gather: "final"  -> profiles at the end of the last slab, dataYs [B, rows, zNo] on every rank (default);
            "all"    -> every slab, dataYs [B, tNo, rows, zNo];
            "outlet" -> last node of the last slab, dataYs [B, rows];
            None     -> nothing but the failure count crosses GPUs.
    rows = y_i..., T [K] (the reference's dataYs rows, pbHomoReactor.py:3660-3677; iso-thermal: y_i only).
    One packed all-gather per call; with keep_on_device the gathered arrays stay device tensors."""
    import torch
    from . import engine
    from .rmt import _check_components
    _check_components(modelInput)
    if modelInput['model'] not in ("N2", "M9"):
        raise NotImplementedError("rmtExeBatchN2Sharded covers the dynamic models N2 and M9")
    if gather not in ("final", "all", "outlet", None):
        raise ValueError("gather must be 'final', 'all', 'outlet' or None")
    B = _batch_size(sweep, B)
    rank, world = world_info(group, comm)
    local, lo, hi = shard_sweep(sweep, B, world, rank)
    Bl = hi - lo
    grid = engine.solverSetting['N2' if modelInput['model'] == "N2" else 'S2']
    zNo = int(grid['zNo'] if zNo is None else zNo)
    tNo = int(grid['tNo'] if tNo is None else tNo)
    ws = workspace if workspace is not None else engine.Workspace()
    dev = torch.device("cuda", torch.cuda.current_device())
    res = None
    if Bl > 0:
        cm = engine.compile_model_n2(modelInput, Bl, zNo)
        res = engine.n2_solve_ensemble(cm, modelInput, local, Bl, zNo=zNo, tNo=tNo, rtol=rtol, atol=atol, out_mode=1,
                                       keep_on_device=True, workspace=ws)
        nrow_out = res.out.shape[1]
    else:
        nrow_out = engine.compile_model_n2(modelInput, 1, zNo).spec.n
    per = {"final": nrow_out*zNo, "all": tNo*nrow_out*zNo, "outlet": nrow_out, None: 0}[gather]
    lay = PackLayout(B, world, per)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        pack = ws.get("d_pack", (lay.length,), torch.float64, device=dev)
        rows, status, tail = lay.views(pack, rank)
        if Bl > 0:
            o = res.out                                                   # [tNo][rows][zNo][Bl]
            if gather == "final":
                rows.view(nrow_out, zNo, Bl).copy_(o[tNo - 1])
            elif gather == "all":
                rows.view(tNo, nrow_out, zNo, Bl).copy_(o)
            elif gather == "outlet":
                rows.copy_(o[tNo - 1, :, zNo - 1, :])
            status.copy_(res.status)
            tail.zero_()
            tail[2] = -1.0
            tail[3] = (res.status != 0).sum()
        else:
            tail.copy_(torch.tensor([0.0, float("inf"), -1.0, 0.0], dtype=torch.float64), non_blocking=True)
        g = all_gather_packed(pack, world, group, comm, stream)
        out = {"range": (lo, hi), "local_out": None if res is None else res.out, "local_status": None if res is None else res.status,
               "local_stats": None if res is None else res.stats, "zNo": zNo, "tNo": tNo,
               "dataTime": np.linspace(0, float(modelInput['operating-conditions']['period']), tNo + 1)[1:],
               "dataXs": np.linspace(0, 1, zNo)}
        shape = {"final": (nrow_out, zNo), "all": (tNo, nrow_out, zNo), "outlet": (nrow_out,), None: ()}[gather]
        if keep_on_device:
            out["gathered"] = g                                            # [world, length] device tensor (PackLayout)
            out["layout"] = lay
            tails = g[:, lay.length - lay.TAIL:].cpu().numpy()
        else:
            full, st, tails = lay.unpack(g.cpu().numpy())
            out["status"] = st
            if gather is not None:
                out["dataYs"] = np.ascontiguousarray(np.moveaxis(full.reshape(shape + (B,)), -1, 0))
        out["failed"] = fold_tails(tails)[3]
    return out
