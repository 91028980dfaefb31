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


_scan_calls = [0]      # every rank calls run_scan the same number of times: a common name for the work counter of a call


def _shared_counter(world):
    """A work counter all ranks can increment atomically: key of the process group's rendezvous store (the TCP store that
    `env://` initialisation creates; no collective, no GPU involved).  Returns ``next_index()`` or ``None`` if the store
    is not usable on EVERY rank (the decision is made collectively, so that all ranks schedule the same way)."""
    key = 'tenpy_b200_scan_%d' % _scan_calls[0]
    store, ok = None, 1
    try:
        store = dist.distributed_c10d._get_default_store()
        store.add(key, 0)
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int64, device=_device(dist.get_backend()))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return None
    return lambda: int(store.add(key, 1)) - 1


def run_scan(configs, run_fn, cost_fn=None, schedule='dynamic'):
    """Run `run_fn(config) -> sequence of floats` for every config, sharded over the ranks.

    `schedule`: ``'dynamic'`` (default) -- the runs are ordered by decreasing estimated cost and every rank pulls the next
    one from a shared counter when it becomes free (list scheduling on the MEASURED run times: the estimate only fixes the
    order); ``'static'`` -- LPT assignment on the estimates before anything runs (:func:`assign_runs`; also the fallback
    when the ranks share no store).  With one rank both are the same loop.
    Returns the gathered (n_runs, 1 + n_values) table (run index first) on every rank."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    costs = [cost_fn(c) if cost_fn else 1. for c in configs]
    _scan_calls[0] += 1
    pull = _shared_counter(world) if (world > 1 and schedule == 'dynamic') else None
    if pull is not None:
        order = sorted(range(len(costs)), key=lambda i: -costs[i])

        def my_runs():
            while True:
                j = pull()
                if j >= len(order):
                    return
                yield order[j]
        mine = my_runs()
    else:
        mine = assign_runs(costs, world)[rank]
    rows = []
    width = None
    failed = []
    for i in mine:
        try:
            vals = [float(v) for v in run_fn(configs[i])]
        except Exception as e:      # keep going: the other ranks wait in the collectives below, a raise here would hang them
            failed.append((i, repr(e)))
            rows.append([float(i)])
            continue
        rows.append([float(i)] + vals)
        width = len(vals) + 1
    n_failed = len(failed)
    if world > 1:
        dev = _device(dist.get_backend())
        w = torch.tensor([width or 0], dtype=torch.int64, device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        width = int(w.item())
        nf = torch.tensor([n_failed], dtype=torch.int64, device=dev)
        dist.all_reduce(nf, op=dist.ReduceOp.SUM)
        n_failed = int(nf.item())
    width = width or 1
    rows = [r + [float('nan')] * (width - len(r)) for r in rows]
    table = gather_results(rows, width)
    if n_failed:
        raise RuntimeError('scan: %d run(s) failed%s' % (n_failed, (': ' + '; '.join('run %d: %s' % f for f in failed)) if failed else
                                                         ' on another rank'))
    return table
