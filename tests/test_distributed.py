"""Host logic of the multi-GPU path on CPU: 2 ranks over gloo.  The rank-local solver is injected (here the CPU
oracle port, tests only); on GPUs it is BatchedRacingMPC.solve and the backend is nccl (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, make_oracle


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import warnings
    warnings.filterwarnings("ignore")
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import racing_lmpc_ros2_b200 as pkg
    from racing_lmpc_ros2_b200 import distributed as D
    from oracle import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    veh = pkg.configs.BARC_VEHICLE; cfg = pkg.configs.barc_lmpc_config(20)
    track = pkg.workload.load_track("barc_center")
    laps = pkg.workload.load_laps() if rank == 0 else []          # only rank 0 owns the laps
    laps = D.broadcast_laps(laps, dist)
    orc = Oracle(veh, cfg)
    for l in laps:
        orc.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 11, 0xD157, track, pkg.workload.load_laps(), mode="barc")   # odd size: uneven shards

    def solve_fn(shard):
        r = orc.step_batch(shard, impl="port", nthreads=1)
        return dict(X_optm=r["X"], U_optm=r["U"], dU_optm=r["dU"], cost=r["cost"], status=r["status"])

    out = D.solve_sharded(solve_fn, batch, cfg["N"], dist)
    q.put((rank, len(laps), out["X_optm"], out["status"], out["cost"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds(pkg):
    from racing_lmpc_ros2_b200.distributed import shard_bounds
    for total in (1, 7, 8, 1024, 65536, 11):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_solve_matches_single_process(pkg):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc")
    batch = pkg.workload.make_batch(veh, cfg, 11, 0xD157, track, pkg.workload.load_laps(), mode="barc")
    ref = o.step_batch(batch, impl="port", nthreads=2)
    for rank, nlaps, X, status, cost in res:
        assert nlaps == 3                                   # rank 1 got the laps through the broadcast
        assert X.shape == (11, 20, 6)
        assert np.array_equal(status, ref["status"])
        assert np.array_equal(X, ref["X"]) and np.array_equal(cost, ref["cost"])   # same code, same inputs: bit-exact


def test_flat_slab_layout_round_trip(pkg):
    """The single-allocation output layout ([X | U | dU | cost | status:int32] per rank) that the multi-GPU bench
    all-gathers without a packing kernel: unpack_flat_slab inverts it, odd batch sizes included."""
    from racing_lmpc_ros2_b200.distributed import unpack_flat_slab
    rng = np.random.default_rng(4)
    for world, Bn, N in ((2, 5, 20), (4, 8, 40), (1, 3, 7)):
        NS = N - 1
        parts, ref = [], dict(X_optm=[], U_optm=[], dU_optm=[], cost=[], status=[])
        for r in range(world):
            X = rng.normal(size=(Bn, N, 6)); U = rng.normal(size=(Bn, NS, 2)); dU = rng.normal(size=(Bn, NS, 2))
            cost = rng.normal(size=Bn); st = rng.integers(0, 5, Bn).astype(np.int32)
            tail = np.zeros((Bn + 1) // 2); tail.view(np.int32)[:Bn] = st
            parts.append(np.concatenate([X.ravel(), U.ravel(), dU.ravel(), cost, tail]))
            for k, v in zip(ref, (X, U, dU, cost, st)):
                ref[k].append(v)
        g = unpack_flat_slab(np.concatenate(parts), world, Bn, N)
        for k in ref:
            assert np.array_equal(g[k], np.concatenate(ref[k])), k


# ---- ShardedSolver (the N-GPU form of the path) on CPU: two ranks over gloo, the rank-local solver replaced by a stand-in
# that fills the output slab from its inputs.  What is tested is the host logic around the exchange: buffer sets, step /
# wait bookkeeping, the slab layout (padded `per`), every rank ending up with every rank's block.  The fused peer-memory
# exchange itself needs GPUs (tests/test_multi_gpu.py).
class _FakeMPC:
    def __init__(self, N):
        self.N, self.K = N, 4

    def _shapes(self, Bn):
        N, K = self.N, self.K
        return dict(X_optm=(Bn, N, 6), U_optm=(Bn, N - 1, 2), dU_optm=(Bn, N - 1, 2), convex_combi_optm=(Bn, K), ss_x=(Bn, K, 6), ss_j=(Bn, K), cost=(Bn,))

    def alloc_device_outputs(self, Bn, device=None):
        import torch
        shp = self._shapes(Bn)
        out = {}
        n64 = sum(int(np.prod(shp[k])) for k in ("X_optm", "U_optm", "dU_optm", "cost"))
        slab = torch.zeros(n64 + (Bn + 1) // 2, dtype=torch.float64)
        o = 0
        for k in ("X_optm", "U_optm", "dU_optm", "cost"):
            n = int(np.prod(shp[k])); out[k] = slab[o:o + n].view(shp[k]); o += n
        out["status"] = slab[o:].view(torch.int32)[:Bn]
        out["slab"] = slab
        out["iters"] = torch.zeros(Bn, dtype=torch.int32)
        return out

    def solve(self, d_in, out):
        s = float(d_in["x_ic"].sum())
        out["X_optm"].copy_(d_in["x_ic"][:, None, :].expand_as(out["X_optm"]) + 1.0)
        out["U_optm"].fill_(s); out["dU_optm"].fill_(-s); out["cost"].copy_(d_in["x_ic"][:, 0] * 2.0)
        out["status"].copy_((d_in["x_ic"][:, 1] > 0).to(out["status"].dtype))
        return out


def _sharded_worker(rank, world, port, q):
    import warnings
    warnings.filterwarnings("ignore")
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from racing_lmpc_ros2_b200 import distributed as D
    from test_distributed import _FakeMPC
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    N, Bn = 7, 5
    res = []
    for backend, sets in (("nccl", 2), ("nccl-overlap", 4)):      # torch.distributed collectives; gloo stands in for NCCL on CPU
        sh = D.ShardedSolver(_FakeMPC(N), dist, Bn, torch.device("cpu"), backend=backend, sets=sets)
        for k in range(5):                                          # more steps than buffer sets: sets are re-used
            g = torch.Generator().manual_seed(100 * k + rank)
            d_in = {"x_ic": torch.randn(Bn, 6, generator=g, dtype=torch.float64)}
            sh.step(d_in, k)
            if backend == "nccl-overlap" and k > 0:
                sh.wait(k - 1)                                      # one step of slack
            if backend == "nccl" or k == 4:
                sh.wait(k)
            if backend == "nccl" or k == 4:
                res.append((backend, k, D.unpack_flat_slab(sh.gathered(k).numpy().ravel().copy(), world, Bn, N, per=sh.per),
                            {key: sh.local(k)[key].numpy().copy() for key in ("X_optm", "cost", "status")}))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solver_host_logic_two_ranks_gloo(pkg):
    import torch
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Bn = 5
    for (b0, k0, g0, own0), (b1, k1, g1, own1) in zip(got[0], got[1]):
        assert (b0, k0) == (b1, k1)
        for key in ("X_optm", "U_optm", "dU_optm", "cost", "status"):
            assert np.array_equal(g0[key], g1[key]), (b0, k0, key)            # both ranks hold the same gathered set
        for r, own in ((0, own0), (1, own1)):                                  # block r is what rank r produced
            for key in ("X_optm", "cost", "status"):
                assert np.array_equal(g0[key][r * Bn:(r + 1) * Bn], own[key]), (b0, k0, r, key)
        x0 = torch.randn(Bn, 6, generator=torch.Generator().manual_seed(100 * k0 + 0), dtype=torch.float64).numpy()
        assert np.array_equal(g0["cost"][:Bn], 2.0 * x0[:, 0])                 # ... from the inputs of that very step
