"""Multi-GPU harness of the path: instances are independent, so the batch is sharded contiguously over the
ranks, the safe set is replicated (rank 0's laps are broadcast once per update), every rank solves its shard,
and ONE collective per batch gathers the trajectories.  The reference has no counterpart (it is a
single process per node); this is the plumbing around `lmpc_solve_batch`, built on torch.distributed
(backend nccl on GPUs; gloo for the CPU tests of this host logic).
"""
import numpy as np


def shard_bounds(total, world, rank):
    """Contiguous, balanced shard [lo, hi) of `total` instances for `rank`."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slab_width(N):
    return 6 * N + 4 * (N - 1) + 2     # X, U, dU, cost, status


def pack_slab(out, N):
    """(b, slab_width) float64 array from a solve() output dict (numpy)."""
    b = out["X_optm"].shape[0]
    s = np.empty((b, slab_width(N)))
    s[:, :6 * N] = out["X_optm"].reshape(b, -1)
    s[:, 6 * N:6 * N + 2 * (N - 1)] = out["U_optm"].reshape(b, -1)
    s[:, 6 * N + 2 * (N - 1):6 * N + 4 * (N - 1)] = out["dU_optm"].reshape(b, -1)
    s[:, -2] = out["cost"]
    s[:, -1] = out["status"]
    return s


def unpack_slab(slab, N):
    b = slab.shape[0]
    return dict(X_optm=slab[:, :6 * N].reshape(b, N, 6), U_optm=slab[:, 6 * N:6 * N + 2 * (N - 1)].reshape(b, N - 1, 2),
                dU_optm=slab[:, 6 * N + 2 * (N - 1):6 * N + 4 * (N - 1)].reshape(b, N - 1, 2), cost=slab[:, -2].copy(),
                status=slab[:, -1].astype(np.int32))


def unpack_flat_slab(flat, world, Bn, N):
    """Inverse of the layout BatchedRacingMPC.alloc_device_outputs gives out["slab"], after an all-gather of it over
    `world` ranks: flat is (world * slab_len,) float64 (numpy).  Returns the global instance-major outputs."""
    NS = N - 1
    n64 = Bn * (6 * N + 4 * NS + 1)
    per = n64 + (Bn + 1) // 2
    flat = np.ascontiguousarray(flat).reshape(world, per)
    X, U, dU, cost, status = [], [], [], [], []
    for r in range(world):
        f = flat[r]
        o = 0
        X.append(f[o:o + Bn * 6 * N].reshape(Bn, N, 6)); o += Bn * 6 * N
        U.append(f[o:o + Bn * 2 * NS].reshape(Bn, NS, 2)); o += Bn * 2 * NS
        dU.append(f[o:o + Bn * 2 * NS].reshape(Bn, NS, 2)); o += Bn * 2 * NS
        cost.append(f[o:o + Bn]); o += Bn
        status.append(f[o:].copy().view(np.int32)[:Bn])
    return dict(X_optm=np.concatenate(X), U_optm=np.concatenate(U), dU_optm=np.concatenate(dU), cost=np.concatenate(cost),
                status=np.concatenate(status))


def broadcast_laps(laps, dist, device=None):
    """Rank 0's laps to every rank (one broadcast of a packed tensor per lap)."""
    import torch
    world = dist.get_world_size()
    if world == 1:
        return laps
    rank = dist.get_rank()
    n = torch.tensor([len(laps) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src=0)
    out = []
    for i in range(int(n.item())):
        m = torch.tensor([laps[i]["x"].shape[0] if rank == 0 else 0], dtype=torch.int64, device=device)
        dist.broadcast(m, src=0)
        rows = int(m.item())
        buf = torch.zeros((rows, 10), dtype=torch.float64, device=device)   # x(6) u(2) k t
        if rank == 0:
            l = laps[i]
            buf.copy_(torch.from_numpy(np.column_stack([l["x"], l["u"], l["k"], l["t"]])))
        dist.broadcast(buf, src=0)
        a = buf.cpu().numpy()
        out.append(dict(x=np.ascontiguousarray(a[:, :6]), u=np.ascontiguousarray(a[:, 6:8]), k=a[:, 8].copy(), t=a[:, 9].copy()))
    return out


def solve_sharded(solve_fn, batch, N, dist=None, device=None):
    """Solve a global batch over all ranks.

    solve_fn(shard_dict) -> output dict (numpy) is the rank-local solver (BatchedRacingMPC.solve on a GPU).
    Every rank passes the same global `batch`; returns the gathered global outputs on every rank.
    """
    import torch
    total = int(batch["x_ic"].shape[0])
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = shard_bounds(total, world, rank)
    out = solve_fn({k: v[lo:hi] for k, v in batch.items()})
    slab = pack_slab(out, N)
    if world == 1:
        return unpack_slab(slab, N)
    # equal-size all-gather: pad the (at most one row shorter) shards
    width = slab.shape[1]
    per = -(-total // world)
    pad = np.zeros((per, width))
    pad[:hi - lo] = slab
    mine = torch.from_numpy(pad).to(device) if device is not None else torch.from_numpy(pad)
    gathered = torch.empty((world * per, width), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine)
    g = gathered.cpu().numpy()
    rows = []
    for r in range(world):
        l, h = shard_bounds(total, world, r)
        rows.append(g[r * per:r * per + (h - l)])
    return unpack_slab(np.vstack(rows), N)
