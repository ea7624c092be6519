"""The N-GPU form of the path on real GPUs (needs >= 2 devices; skipped otherwise): two ranks over NCCL, the
trajectory exchange fused into the QP kernel over NVLink peer memory (lmpc_solve_gather_batch) against the plain
ncclAllGather of the same results.  Bit-exact: the exchange moves bytes."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import warnings
    warnings.filterwarnings("ignore")
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import racing_lmpc_ros2_b200 as pkg
    from racing_lmpc_ros2_b200 import distributed as D
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    veh = pkg.configs.BARC_VEHICLE; cfg = pkg.configs.barc_lmpc_config(20)
    track = pkg.workload.load_track("barc_center"); laps = pkg.workload.load_laps()
    Bn, N = 96, cfg["N"]
    res = {}
    for backend in ("peer", "nccl"):
        mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn, device=rank)
        for l in laps:
            mpc.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
        sh = D.ShardedSolver(mpc, dist, Bn, dev, backend=backend)
        outs = []
        for k in range(3):      # three steps: both buffer sets are used and one is re-used
            data = pkg.workload.make_batch(veh, cfg, Bn, 0x6A7 + 100 * k + rank, track, laps, mode="barc")
            d_in = {key: torch.from_numpy(v).to(dev) for key, v in data.items()}
            sh.step(d_in, k); sh.wait(k)
            torch.cuda.synchronize(dev)
            g = D.unpack_flat_slab(sh.gathered(k).cpu().numpy(), world, Bn, N, per=sh.per)
            own = {key: sh.local(k)[key].cpu().numpy() for key in ("X_optm", "U_optm", "dU_optm", "cost", "status")}
            outs.append((g, own))
        if backend == "peer":
            assert mpc.gather_error() == 0
            # the host-buffer form: H2D, solve + fused exchange, wait, D2H of the whole gathered set
            h_in = mpc.alloc_host_inputs(data, pinned=True)
            h_out = mpc.alloc_host_outputs(Bn, pinned=True)
            g_host = np.zeros(world * sh.per)
            mpc.set_stream(None)
            mpc.solve_gather(h_in, h_out, 1, wait=True, gathered_host=g_host)
            gh = D.unpack_flat_slab(g_host, world, Bn, N, per=sh.per)
            res["host"] = gh
        dist.barrier()
        res[backend] = outs
        mpc.close()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_fused_peer_exchange_equals_nccl_all_gather():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for k in range(3):
        for key in ("X_optm", "U_optm", "dU_optm", "cost", "status"):
            a0, a1 = got[0]["peer"][k][0][key], got[1]["peer"][k][0][key]
            n0 = got[0]["nccl"][k][0][key]
            assert np.array_equal(a0, a1), (k, key)          # both ranks hold the same gathered set
            assert np.array_equal(a0, n0), (k, key)          # ... equal to the ncclAllGather of the same solves
            for r in range(world):                           # ... whose block r is what rank r solved
                own = got[r]["peer"][k][1][key]
                assert np.array_equal(a0[r * 96:(r + 1) * 96], own), (k, key, r)
        assert (got[0]["peer"][k][0]["status"] == 0).mean() > 0.98
    for key in ("X_optm", "status"):
        assert np.array_equal(got[0]["host"][key], got[0]["peer"][2][0][key]), key
        assert np.array_equal(got[1]["host"][key], got[0]["host"][key]), key


@pytest.mark.gpu
def test_solve_gather_on_one_gpu_equals_plain_solve(pkg):
    """world = 1: lmpc_solve_gather_batch routes the trajectory outputs into the (one-rank) gathered set and mirrors to
    nobody; device and host forms must reproduce lmpc_solve_batch bit for bit, and the set layout must unpack."""
    import torch
    from racing_lmpc_ros2_b200 import distributed as D
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    veh = pkg.configs.BARC_VEHICLE; cfg = pkg.configs.barc_lmpc_config(20)
    track = pkg.workload.load_track("barc_center"); laps = pkg.workload.load_laps()
    Bn, N = 48, cfg["N"]
    mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn)
    for l in laps:
        mpc.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    data = pkg.workload.make_batch(veh, cfg, Bn, 0x1C, track, laps, mode="barc")
    plain = mpc.solve(data)
    dev = torch.device("cuda", 0)
    sh = D.ShardedSolver(mpc, None, Bn, dev, backend="peer", sets=2)
    d_in = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    for k in range(3):
        sh.step(d_in, k); sh.wait(k)
    torch.cuda.synchronize()
    assert mpc.gather_error() == 0
    g = D.unpack_flat_slab(sh.gathered(2).cpu().numpy(), 1, Bn, N, per=sh.per)
    for key in ("X_optm", "U_optm", "dU_optm", "cost", "status"):
        assert np.array_equal(g[key], plain[key]), key
        assert np.array_equal(sh.local(2)[key].cpu().numpy(), plain[key]), key
    assert np.array_equal(sh.local(2)["iters"].cpu().numpy(), plain["iters"])
    # host form: own shard back through the caller's buffers, or the whole set through gathered_host
    h_in = mpc.alloc_host_inputs(data, pinned=True); h_out = mpc.alloc_host_outputs(Bn, pinned=True)
    mpc.solve_gather(h_in, h_out, 0, wait=True)
    for key in ("X_optm", "U_optm", "dU_optm", "cost", "status", "iters", "convex_combi_optm", "ss_x", "ss_j"):
        assert np.array_equal(h_out[key], plain[key]), key
    g_host = np.zeros(sh.per)
    mpc.solve_gather(h_in, h_out, 1, wait=True, gathered_host=g_host)
    gh = D.unpack_flat_slab(g_host, 1, Bn, N, per=sh.per)
    assert np.array_equal(gh["X_optm"], plain["X_optm"]) and np.array_equal(gh["status"], plain["status"])
