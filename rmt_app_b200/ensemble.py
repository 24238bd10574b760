"""Ensemble sharding across the GPUs of one box (one process per GPU).

Reactor instances are independent initial-value problems — nothing couples them
(the reference solves exactly one per `rmtExe` call) — so the ensemble is split
into contiguous index blocks, one per rank, and there is NO collective on the data
path.  `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) is used
only at the end of a run:

* parameter-estimation populations: all-reduce of the objective's (sum, min, argmin),
  each rank contributing the three numbers its GPU reduced from its own shard
  (`rmt_reduce_objective`), plus an optional all-gather of the per-instance
  objectives (8 B each);
* result gathering: all-gather (or gather to rank 0) of outlet rows.

The helpers below only move tensors; they run unchanged on CPU tensors under gloo,
which is how the host logic is tested without a GPU (tests/test_distributed_cpu.py).
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


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def reduce_objective(local_sum, local_min, local_argmin, device=None, group=None):
    """Global (sum, min, argmin) of a sharded objective from the per-rank triples.
    Ties on the minimum resolve to the smallest global index, so the result does not
    depend on the number of ranks."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return float(local_sum), float(local_min), int(local_argmin)
    t = torch.tensor([float(local_sum), float(local_min), float(local_argmin)], dtype=torch.float64, device=device)
    allt = torch.empty((world, 3), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allt.view(-1), t, group=group)
    allt = allt.cpu().numpy()
    total = float(np.sum(allt[:, 0]))            # fixed rank order: deterministic
    mn = allt[:, 1].min()
    cand = allt[allt[:, 1] == mn, 2]
    return total, float(mn), int(cand.min())


def all_gather_rows(local, B, group=None):
    """Concatenate per-rank shards along the LAST dimension (the instance index) into
    the full ensemble on every rank.  `local`: tensor [..., B_local]."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    sizes = [partition(B, world, r)[1] - partition(B, world, r)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros(local.shape[:-1] + (mx,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    out = torch.empty((world*pad.numel(),), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous().view(-1), group=group)
    out = out.view((world,) + tuple(pad.shape))
    return torch.cat([out[r][..., :sizes[r]] for r in range(world)], dim=-1)


def rmtExeBatchSharded(modelInput, sweep, B=None, *, rtol=None, atol=None, objective_ref=None, gather=True,
                       group=None, workspace=None):
    """Every rank calls this with the SAME full `sweep`; each solves its block on its own GPU.

    Returns a dict with the rank's shard ("local_*", device tensors), the global objective
    statistics when `objective_ref` is given, and — with `gather=True` — the full outlet
    array [B, n] (host numpy) on every rank."""
    import torch
    from . import engine
    from .rmt import _check_components
    _check_components(modelInput)
    if B is None:
        first = next(iter(sweep.values()))
        B = int(first.shape[0]) if hasattr(first, "shape") else len(first)
    rank, world = world_info(group)
    local, lo, hi = shard_sweep(sweep, B, world, rank)
    rt = modelInput.get('solver-config', {}).get('rtol', engine.DEFAULT_RTOL) if rtol is None else rtol
    cm = engine.compile_model(modelInput, method=engine.choose_method(modelInput, rt, 1))
    res = engine.n1_solve_ensemble(cm, modelInput, local, hi - lo, rtol=rtol, atol=atol, out_mode=1,
                                   objective_ref=objective_ref, keep_on_device=True, workspace=workspace)
    out = {"range": (lo, hi), "local_dataYs": res.out[0], "local_status": res.status, "local_stats": res.stats}
    dev = res.out.device
    if objective_ref is not None:
        s, mn, am = cm.module.reduce_objective(hi - lo, res.objective, index_offset=lo,
                                               stream=torch.cuda.current_stream().cuda_stream)
        out["objective_sum"], out["objective_min"], out["objective_argmin"] = reduce_objective(s, mn, am, dev, group)
        out["local_objective"] = res.objective
        if gather:
            out["objective"] = all_gather_rows(res.objective, B, group).cpu().numpy()
    nbad = torch.tensor([float((res.status != 0).sum().item())], dtype=torch.float64, device=dev)
    if world > 1:
        _dist().all_reduce(nbad, group=group)
    out["failed"] = int(nbad.item())
    if gather:
        full = all_gather_rows(res.out[0], B, group)                 # [n][B]
        out["dataYs"] = full.t().contiguous().cpu().numpy()
        out["status"] = all_gather_rows(res.status, B, group).cpu().numpy()
    return out
