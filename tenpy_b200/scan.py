"""Sharding of independent DMRG runs (a parameter scan over bond dimension chi / field g) over ranks.

A single DMRG run is a strict site-to-site dependency chain (reference mps_common.py:394-405) and does not
shard; independent runs do (BASELINE.json configs[4]).  One process per GPU; the only
collectives are a broadcast of the model template from rank 0 and an all-gather of the per-run results --
there is no collective on the data path.  Works with the `nccl` backend (GPU tensors) and with `gloo` (CPU
tensors; used by the CPU tests).
"""
# Copyright (C) 2026 tenpy_b200 authors. Apache-2.0.

import numpy as np
import torch
import torch.distributed as dist

__all__ = ['assign_runs', 'broadcast_template', 'gather_results', 'run_scan']


def assign_runs(costs, world_size):
    """Assign runs to ranks: largest estimated cost first, always to the least loaded rank (LPT).

    `costs[i]` is the cost estimate of run `i` (e.g. chi**3).  Returns a list of index lists, one per rank."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(i)
        load[r] += float(costs[i])
    return out


def _device(backend):
    return torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')


def broadcast_template(values, src=0):
    """Broadcast a small float64 vector (the model template: couplings, sizes) from rank `src` to all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(values, dtype=np.float64)
    t = torch.tensor(np.asarray(values, dtype=np.float64), device=_device(dist.get_backend()))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def gather_results(rows, width):
    """All-gather per-run result rows (each `width` float64 numbers, first entry = run index).

    Ranks may own different numbers of runs; rows are padded with NaN.  Returns an (n_runs, width) array sorted
    by run index (on every rank)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, width)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows[np.argsort(rows[:, 0])] if len(rows) else rows
    world = dist.get_world_size()
    dev = _device(dist.get_backend())
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = max(int(c.item()) for c in counts)
    pad = np.full((nmax, width), np.nan)
    pad[:rows.shape[0]] = rows
    mine = torch.tensor(pad, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    res = np.concatenate([a.cpu().numpy()[:int(c.item())] for a, c in zip(allr, counts)], axis=0)
    return res[np.argsort(res[:, 0])]


def run_scan(configs, run_fn, cost_fn=None):
    """Run `run_fn(config) -> sequence of floats` for every config, sharded over the ranks.

    Returns the gathered (n_runs, 1 + n_values) table (run index first) on every rank."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    costs = [cost_fn(c) if cost_fn else 1. for c in configs]
    mine = assign_runs(costs, world)[rank]
    rows = []
    width = None
    for i in mine:
        vals = [float(v) for v in run_fn(configs[i])]
        rows.append([float(i)] + vals)
        width = len(vals) + 1
    if world > 1:
        w = torch.tensor([width or 0], dtype=torch.int64, device=_device(dist.get_backend()))
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        width = int(w.item())
    return gather_results(rows, width or 1)
