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


def unpack_flat_slab(flat, world, Bn, N, per=None):
    """Inverse of the layout BatchedRacingMPC.alloc_device_outputs gives out["slab"] (and lmpc_gather_init gives one
    rank's block, there padded to `per` doubles), after a gather over `world` ranks: flat is (world * per,) float64
    (numpy).  Returns the global instance-major outputs."""
    NS = N - 1
    n64 = Bn * (6 * N + 4 * NS + 1)
    per = per or (n64 + (Bn + 1) // 2)
    flat = np.ascontiguousarray(flat).reshape(world, per)
    X, U, dU, cost, status = [], [], [], [], []
    for r in range(world):
        f = flat[r]
        o = 0
        X.append(f[o:o + Bn * 6 * N].reshape(Bn, N, 6)); o += Bn * 6 * N
        U.append(f[o:o + Bn * 2 * NS].reshape(Bn, NS, 2)); o += Bn * 2 * NS
        dU.append(f[o:o + Bn * 2 * NS].reshape(Bn, NS, 2)); o += Bn * 2 * NS
        cost.append(f[o:o + Bn]); o += Bn
        status.append(f[o:o + (Bn + 1) // 2].copy().view(np.int32)[:Bn])
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


class _DevPtr:
    """A raw device allocation as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


class ShardedSolver:
    """The N-GPU form of the path (one process per GPU): this rank solves its shard of the batch and every rank ends
    up with the trajectories (X, U, dU, cost, status) of ALL ranks -- the one exchange SURVEY.md 8e names.

    backend "peer" (default): the exchange is fused into the QP kernel -- its epilogue stores every instance's result
        into all peers' gather buffers over NVLink peer mappings (lmpc_solve_gather_batch); nothing but a one-warp wait
        kernel runs for the collective, and no communication kernel sits on the SMs beside the solve.
    backend "nccl": solve into a local slab, then ncclAllGather of the slab (torch.distributed), stream-ordered after the
        solve.  "nccl-overlap": the same, issued asynchronously so that it runs beside the NEXT solve (round 1's scheme,
        kept for A/B: the resident NCCL kernel takes SM slots from a one-wave QP grid).
    Every rank must use the same per-rank batch size.  step(k) enqueues solve k and its exchange (into buffer set
    k % sets); wait(k) makes gathered(k) valid in stream order.  sets = 2 is safe when wait(k) is enqueued before
    step(k + 1).  Running one step ahead -- step(k + 1) before wait(k), so that a rank does not idle while a slower peer
    finishes -- needs sets = 4: a consumer of gathered(k) enqueued right after wait(k) is then ordered before this rank's
    step(k + 2), whose completion is what allows a peer to start step(k + 4) and overwrite that set.
    """

    def __init__(self, mpc, dist, Bn, device, backend="peer", sets=2):
        import torch
        self.mpc, self.dist, self.Bn, self.dev, self.backend, self.sets = mpc, dist, int(Bn), device, backend, int(sets)
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.N = mpc.N
        self._seq = {}
        self._pending = {}
        if backend == "peer":
            handle, self.per = mpc.gather_init(self.world, self.rank, self.Bn, self.sets)
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, handle)
                mpc.gather_connect(handles)
                dist.barrier()    # every rank has mapped every peer before the first remote store
            self.gbuf = []
            for s in range(self.sets):
                ptr, per = mpc.gather_buffer_ptr(s)
                self.gbuf.append(torch.as_tensor(_DevPtr(ptr, per * self.world), device=device))
            shp = mpc._shapes(self.Bn)
            self.outs = [{k: torch.empty(shp[k], dtype=torch.float64, device=device) for k in ("convex_combi_optm", "ss_x", "ss_j")}
                         for _ in range(self.sets)]
            for o in self.outs:
                o["iters"] = torch.empty(self.Bn, dtype=torch.int32, device=device)
        elif backend in ("nccl", "nccl-overlap"):
            self.outs = [mpc.alloc_device_outputs(self.Bn, device) for _ in range(self.sets)]
            self.per = self.outs[0]["slab"].numel()
            self.gbuf = [torch.empty(self.world * self.per, dtype=torch.float64, device=device) for _ in range(self.sets)]
        else:
            raise ValueError(backend)

    def step(self, d_in, k):
        s = k % self.sets
        if self.backend == "peer":
            self._seq[k] = self.mpc.solve_gather(d_in, self.outs[s], s, wait=False)
        else:
            self.mpc.solve(d_in, self.outs[s])
            if self.world > 1:
                h = self.dist.all_gather_into_tensor(self.gbuf[s], self.outs[s]["slab"], async_op=(self.backend == "nccl-overlap"))
                if self.backend == "nccl-overlap":
                    self._pending[k] = h
            else:
                self.gbuf[s].copy_(self.outs[s]["slab"])

    def wait(self, k):
        if self.backend == "peer":
            seq = self._seq.pop(k, None)
            if seq is not None:
                self.mpc.gather_wait(seq)
        else:
            h = self._pending.pop(k, None)
            if h is not None:
                h.wait()

    def gathered(self, k):
        """(world, per) view of set k % sets: row r is rank r's slab (distributed.unpack_flat_slab(.., per=self.per))."""
        return self.gbuf[k % self.sets].view(self.world, self.per)

    def local(self, k):
        """This rank's own outputs of step k as tensors (views into the gathered set for the trajectory keys)."""
        s = k % self.sets
        if self.backend != "peer":
            return self.outs[s]
        import torch
        N, NS, Bn = self.N, self.N - 1, self.Bn
        row = self.gathered(k)[self.rank]
        o = dict(self.outs[s])
        a = 0
        o["X_optm"] = row[a:a + Bn * 6 * N].view(Bn, N, 6); a += Bn * 6 * N
        o["U_optm"] = row[a:a + Bn * 2 * NS].view(Bn, NS, 2); a += Bn * 2 * NS
        o["dU_optm"] = row[a:a + Bn * 2 * NS].view(Bn, NS, 2); a += Bn * 2 * NS
        o["cost"] = row[a:a + Bn]; a += Bn
        o["status"] = row[a:a + (Bn + 1) // 2].view(torch.int32)[:Bn]
        return o
