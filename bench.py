#!/usr/bin/env python
"""bench.py -- batched LMPC steps/s on N B200s (BASELINE.json's metric) + roofline + CPU baseline.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU arm: the oracle port on all host threads

One *step* = one pass of the hot path (abscissa alignment -> linearisation -> safe-set query -> QP ->
outputs, i.e. RacingMPC::solve) over one batch of synthetic ticks.  Workload = BASELINE.json
configs[1]: BARC LMPC, N=20, 6-state Frenet bicycle, 1024 random initial states per GPU, the three
recorded laps as the safe set (K=96).  Weak scaling: every rank solves its own 1024 instances and the
ranks all-gather the trajectories (one NCCL collective per step, inside the timed region).

value  : whole-job steps/s with inputs resident in HBM, CUDA events on the launching stream, L2 flushed
         (untimed) between timed steps, max over ranks.
e2e    : the same metric through the C-ABI call with HOST (pinned) buffers: H2D + kernels + D2H per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; the contract is ONE JSON line there.  Must be set before
# torch (and with it NCCL) is loaded.  An explicit INFO / TRACE request is left alone.
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "batched LMPC steps/sec (N=20, 6-state bicycle)"
UNIT = "steps/s"
PER_GPU_BATCH = 1024
N_HORIZON = 20


def b_alg(N, K, lam_returned=True):
    """Algorithmic bytes of one step: every input read once, every output written once."""
    b = 8 * ((8 + 6 * N + 2 * (N - 1) + (N - 1) + 4 * N + 1) + (6 * N + 4 * (N - 1) + 2))
    return b + (8 * K if lam_returned else 0)


def workload(pkg, seed, batch):
    veh = pkg.configs.BARC_VEHICLE
    cfg = pkg.configs.barc_lmpc_config(N_HORIZON)
    track = pkg.workload.load_track("barc_center")
    laps = pkg.workload.load_laps()
    data = pkg.workload.make_batch(veh, cfg, batch, seed, track, laps, mode="barc")
    return veh, cfg, track, laps, data


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML in a thread every 2 ms (the region is
    only tens of milliseconds long; `nvidia-smi -lms` would get one sample or none), nvidia-smi as the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index = index
        self.sm, self.mask, self.max_mhz = [], 0, None
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            f = [x.strip() for x in o.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi, one query right after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}

    def stop(self):
        if self.thread is None:
            return self._smi_once()
        self.stop_flag.set()
        self.thread.join(timeout=1.0)
        if not self.sm:
            return self._smi_once()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm),
                "source": "NVML, every 2 ms during the timed region"}


def cpu_arm(pkg, steps, warmup, sample_instances, seed=0xC0DE):
    """The CPU implementation of the path: oracle port (the reference's CasADi/OSQP stack cannot be
    built here), all host threads, on a bounded sample of the same workload."""
    from oracle import Oracle
    veh, cfg, track, laps, data = workload(pkg, seed, sample_instances)
    orc = Oracle(veh, cfg)
    for l in laps:
        orc.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    cores = os.cpu_count() or 1
    for _ in range(warmup):
        orc.step_batch(data, impl="port", nthreads=cores)
    t0 = time.perf_counter()
    nfail = 0
    for _ in range(steps):
        r = orc.step_batch(data, impl="port", nthreads=cores)
        nfail += int(r["nfail"])
    dt = time.perf_counter() - t0
    return dict(value=steps * sample_instances / dt, seconds=dt, cores=cores, failed=nfail,
                sample=f"{steps} steps x {sample_instances} instances of the same workload (oracle port, {cores} threads; {warmup} untimed warm-up steps)")


def probe_reference_stack():
    """SURVEY.md 8c / BASELINE.md 2: never assume the real reference stack is absent -- look for it at run time."""
    found = {"casadi": False, "osqp": False, "baseline_ref": os.path.isdir(os.path.join(ROOT, "baseline", "_ref")),
             "oracle_ref": os.path.isdir(os.path.join(ROOT, "oracle", "_ref"))}
    for mod in ("casadi", "osqp"):
        try:
            __import__(mod)
            found[mod] = True
        except Exception:
            pass
    found["usable"] = bool(found["casadi"])
    found["note"] = ("CasADi importable: oracle/casadi_reference.py can state the reference's Opti('conic')+OSQP problem literally"
                     if found["casadi"] else
                     "CasADi / OSQP not importable and no baseline/_ref: the reference's solver stack is absent on this box; the CPU arm is the oracle port (kind 'port')")
    return found


def timed_steps(stream, flush, steps, body):
    """`steps` repetitions of body(k), each bracketed by CUDA events on `stream`, the L2 flushed (untimed) before each."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        for k in range(steps):
            flush.fill_(float(k))            # untimed L2 flush
            ev[k][0].record(stream)
            body(k)
            ev[k][1].record(stream)
    stream.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def extra_config_line(pkg, name, dev, local_rank, stream, flush, dist, world, rank, peak, steps=5, warmup=3):
    """One line of the `configs` object: BASELINE configs[2..4] at the per-GPU share they name, device-resident inputs,
    CUDA events, L2 flushed between steps, max over ranks.  Returns a dict (rank 0) or None."""
    import torch
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    from racing_lmpc_ros2_b200.distributed import ShardedSolver
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    laps = pkg.workload.load_laps()
    exchange = False
    if name == "config3_iac_tracking_N40":
        veh, cfg = pkg.configs.IAC_VEHICLE, pkg.configs.iac_tracking_config(40)
        track = pkg.workload.load_track("putnam_optm")
        Bn, named_gpus, mode, use_laps, spec = 4096, 1, "track", [], None
        desc = "BASELINE configs[2]: IAC Putnam full-course tracking MPC, N=40, 4096 instances per GPU"
    elif name == "config4_50lap_regression":
        veh = pkg.configs.BARC_VEHICLE
        # iteration cap 60: with the learned error dynamics the linear rollouts start further from the optimum (the car's yaw
        # model error is 0.13 rad/s per step in the recorded laps); 5.6 % of these QPs need more than the default 30
        cfg = dict(pkg.configs.barc_lmpc_config(20), num_ss_pts_per_lap=2, max_lap_stored=50, max_iter=60)
        track = pkg.workload.load_track("barc_center")
        use_laps = pkg.workload.synthesise_laps(laps, 50)
        spec = make_reg_spec([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]], 0.6)
        Bn, named_gpus, mode = 2048, 4, "barc"
        desc = ("BASELINE configs[3]: BARC LMPC, 50-lap safe set (~66 k points; 2 nearest per lap over 48 laps = 96 columns), "
                "error-dynamics regression on every stage (lmpc_set_error_dynamics), 8192 instances on 4 GPUs = 2048 per GPU")
    else:
        veh, cfg = pkg.configs.BARC_VEHICLE, pkg.configs.barc_lmpc_config(20)
        track = pkg.workload.load_track("barc_center")
        use_laps, spec = laps, None
        Bn, named_gpus, mode, exchange = 8192, 8, "barc", True
        desc = "BASELINE configs[4]: Monte-Carlo LMPC, 65536 perturbed agents, N=20, sharded over 8 GPUs = 8192 per GPU, trajectories of all ranks exchanged every step"
    mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn, device=local_rank)
    for l in use_laps:
        mpc.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    if spec is not None:
        mpc.set_error_dynamics(spec)
    data = pkg.workload.make_batch(veh, cfg, Bn, 0xB200 + 40 + 7919 * rank, track, laps, mode=mode)
    d_in = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    mpc.set_stream(stream)
    sh = ShardedSolver(mpc, dist if world > 1 else None, Bn, dev, backend="peer") if exchange else None
    out = mpc.alloc_device_outputs(Bn, dev) if not exchange else None

    def body(k):
        if exchange:
            sh.step(d_in, k); sh.wait(k)
        else:
            mpc.solve(d_in, out)

    with torch.cuda.stream(stream):
        for k in range(warmup):
            body(k)
    stream.synchronize()
    if world > 1:
        dist.barrier()
    mpc.set_timing(True)
    ms = timed_steps(stream, flush, steps, body)
    (ms_lin, ms_ss, ms_qp), nrec = mpc.kernel_ms()
    mpc.set_timing(False)
    t = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    o = sh.local(steps - 1) if exchange else out
    status = o["status"].cpu().numpy(); iters = o["iters"].cpu().numpy()
    if exchange:
        assert mpc.gather_error() == 0
    mpc.close()
    N, K = cfg["N"], (cfg["num_ss_pts"] if cfg["learning"] else 0)
    qp_ms = ms_qp / max(nrec, 1)
    total_ms = float(t.item())
    summ = load_summary(name)
    flops = summ.get("fp64_flops_per_launch") if summ.get("batch") == Bn else None
    line = {"workload": desc, "n_gpus": world, "per_gpu_batch": Bn, "global_batch": Bn * world,
            "is_the_named_configuration": world == named_gpus, "named_n_gpus": named_gpus,
            "value": Bn * world * steps / (total_ms * 1e-3), "unit": UNIT, "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
            "solved_fraction": float(((status == 0) | (status == 5)).mean()), "solved_exact_fraction": float((status == 0).mean()),
            "status_histogram": np.bincount(status, minlength=7).tolist(),
            "ipm_iters_mean": float(iters.mean()), "ipm_iters_max": int(iters.max()),
            "kernel_ms": {"lmpc_qp_kernel": qp_ms, "linearise_and_regression": ms_lin / max(nrec, 1), "ss_query_wait": ms_ss / max(nrec, 1)},
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_step": b_alg(N, K, bool(K)), "achieved": b_alg(N, K, bool(K)) * Bn / (qp_ms * 1e-3) / 1e9 if qp_ms > 0 else None,
                         "peak": peak, "unit": "GB/s", "frac": (b_alg(N, K, bool(K)) * Bn / (qp_ms * 1e-3) / 1e9 / peak) if qp_ms > 0 else None,
                         "fp64_flops_per_launch": flops}}
    return line if rank == 0 else None


def load_summary(name=None):
    path = os.path.join(ROOT, "profiles", "qp_kernel_summary.json" if not name else f"qp_kernel_summary_{name}.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="instances per GPU (default: BASELINE config 2)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl", "nccl-overlap"],
                    help="how the ranks exchange trajectories: fused into the QP kernel over NVLink peer memory (default), or ncclAllGather after / beside the solve (A/B)")
    ap.add_argument("--inputs", default="pool", choices=["pool", "fixed"],
                    help="pool: a different random batch every timed step, one contiguous timed region (default); fixed: one batch, L2 flushed (untimed) between steps")
    ap.add_argument("--sync", default="slack", choices=["slack", "barrier"],
                    help="slack: a rank waits for its peers' results of the PREVIOUS step after enqueuing its next one (two buffer sets); barrier: every step ends with the wait for its own exchange")
    ap.add_argument("--same-seed", action="store_true", help="A/B aid: every rank solves rank 0's batch (isolates the max-over-ranks effect)")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[2..4] lines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import racing_lmpc_ros2_b200 as pkg

    exchange_desc = {"peer": "trajectories of all ranks exchanged every step by the QP kernel's epilogue (stores into every peer's gather buffer over NVLink peer memory, sequence flag, one-warp wait kernel); solve + exchange + wait inside every timed window",
                     "nccl": "one ncclAllGather of the result slab per step, stream-ordered after the solve, inside the timed window",
                     "nccl-overlap": "one ncclAllGather per step issued asynchronously beside the next solve and waited inside that step's window (round 1's scheme)"}[args.gather]
    config_desc = {"workload": f"BASELINE configs[1]: BARC LMPC, N={N_HORIZON}, 6-state Frenet bicycle, K=96 safe-set columns "
                               f"(3 recorded laps), {args.batch} random initial states per GPU",
                   "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1), "N": N_HORIZON, "K": 96,
                   "parallelism": f"instances sharded over {max(world, 1)} GPU(s), safe set replicated; {exchange_desc}",
                   "gather": args.gather, "same_seed_on_all_ranks": bool(args.same_seed),
                   "inputs": args.inputs, "sync": args.sync,
                   "l2": ("a different random batch every timed step (pool of min(64, steps + warmup) batches per GPU, batch j = seed + 104729 j; "
                          "the warm-up uses the other end of the pool), the L2 flushed once before the ONE contiguous timed region: no step finds its inputs in the L2"
                          if args.inputs == "pool" else "256 MiB scratch written between timed steps (untimed) to flush the 126 MB L2"), "tol": 1e-7, "polish": "active-set (augmented-Lagrangian) polish after the interior point"}
    ref_stack = probe_reference_stack()

    # ------------------------------------------------------------------ CPU ("reference") arm
    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.batch * max(args.gpus, 1)     # one step = the arm's global batch
        res = cpu_arm(pkg, steps=max(1, min(args.steps, 20)), warmup=max(1, min(args.warmup, 3)), sample_instances=sample)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / res["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_desc,
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
                                 "failed_instances": res["failed"],
                                 "note": "CPU restatement (oracle/oracle_port.c, same algorithm as the kernel) -- the reference's CasADi/OSQP stack was probed for at run time and is absent", "reference_stack_probe": ref_stack},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    from racing_lmpc_ros2_b200.distributed import ShardedSolver, unpack_flat_slab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solve path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    seed = 0xB200 + 2 + (0 if args.same_seed else 7919 * rank)
    veh, cfg, track, laps, data = workload(pkg, seed, args.batch)
    mpc = BatchedRacingMPC(veh, cfg, max_batch=args.batch, device=local_rank)
    # safe set: rank 0 owns the laps, every rank receives them (replicated), then ingests locally
    for l in laps:
        x = torch.from_numpy(np.ascontiguousarray(l["x"])).to(dev)
        if world > 1:
            dist.broadcast(x, src=0)
        mpc.add_lap(x.cpu().numpy(), l["u"], l["k"], l["t"], track["length"])

    stream = torch.cuda.Stream(device=dev)
    mpc.set_stream(stream)
    N, K = cfg["N"], cfg["num_ss_pts"]
    # Inputs: "pool" (default) = a different random batch every step (seed + 104729 * j), so no step finds its inputs in
    # the L2 and no single slow instance is solved over and over; "fixed" = round 1's scheme (one batch, the L2 flushed
    # between steps).  Batch 0 of the pool is the round-1 batch (the one profiles/qp_kernel_summary.json was captured on).
    n_pool = 1 if args.inputs == "fixed" else min(64, args.steps + args.warmup)
    pool = [data] + [pkg.workload.make_batch(veh, cfg, args.batch, seed + 104729 * j, track, laps, mode="barc") for j in range(1, n_pool)]
    d_pool = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in pool]
    slack = args.sync == "slack" and args.inputs == "pool"
    # buffer sets: 2 when every step ends with its own wait; 4 with one step of slack (a consumer of set k enqueued after
    # wait(k) is then stream-ordered before the solve whose completion lets a peer overwrite that set)
    sharded = ShardedSolver(mpc, dist if world > 1 else None, args.batch, dev, backend=args.gather, sets=4 if slack else 2)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def batch_of(k):          # timed step k -> pool index; the warm-up steps use the END of the pool
        return d_pool[k % n_pool]

    with torch.cuda.stream(stream):
        for j in range(args.warmup):
            k = n_pool - 1 - j if n_pool > 1 else j
            sharded.step(d_pool[k % n_pool], j); sharded.wait(j)
    stream.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    mpc.set_timing(True)
    launches0 = mpc.launch_count
    K0 = 4 * ((args.warmup + 3) // 4)       # step numbering continues after the warm-up (buffer set = step % sets)
    with torch.cuda.stream(stream):
        flush.fill_(1.0)                     # untimed: nothing the timed steps read is left in the L2
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if n_pool > 1:
        # ONE contiguous timed region: K steps on K different batches, every exchange and every wait inside it
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for k in range(args.steps):
                sharded.step(batch_of(k), K0 + k)
                if not slack:
                    sharded.wait(K0 + k)
                else:
                    if k > 0:
                        sharded.wait(K0 + k - 1)
                    if k == args.steps - 1:
                        sharded.wait(K0 + k)
            ev1.record(stream)
        stream.synchronize()
        dev_ms = ev0.elapsed_time(ev1)
    else:
        # fixed inputs (round 1's scheme): the L2 is flushed (untimed) between steps, each step timed on its own and ended
        # by the wait for its own exchange
        def fixed_step(k):
            sharded.step(d_pool[0], K0 + k)
            sharded.wait(K0 + k)
        dev_ms = sum(timed_steps(stream, flush, args.steps, fixed_step))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    launches = mpc.launch_count - launches0
    (ms_lin, ms_ss, ms_qp), nrec = mpc.kernel_ms()
    mpc.set_timing(False)
    clocks = sampler.stop()
    own_ms = dev_ms
    qp_ms = ms_qp / max(nrec, 1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    per_rank = torch.tensor([qp_ms, own_ms / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allr = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(allr, per_rank)
        per_rank_qp = [float(a[0]) for a in allr]; per_rank_step = [float(a[1]) for a in allr]
    else:
        per_rank_qp = [qp_ms]; per_rank_step = [own_ms / args.steps]
    dev_ms = float(t.item())
    last = K0 + args.steps - 1
    last_batch = pool[(args.steps - 1) % n_pool]
    lo_ = sharded.local(last)
    status = lo_["status"].cpu().numpy()
    iters = lo_["iters"].cpu().numpy()
    X_own = lo_["X_optm"].cpu().numpy()
    # the profiled batch (pool batch 0 = round 1's batch) on its own, untimed for the headline: its kernel time is what the
    # fp64 roofline divides the ncu-counted flops of that same launch by; its outputs are what the e2e leg must reproduce
    mpc.set_timing(True)
    with torch.cuda.stream(stream):
        for _ in range(3):
            ref0 = mpc.solve(d_pool[0])
    stream.synchronize()
    (_, _, ms_qp0), nrec0 = mpc.kernel_ms()
    mpc.set_timing(False)
    qp_ms_batch0 = ms_qp0 / max(nrec0, 1)
    status0 = ref0["status"].cpu().numpy(); X0 = ref0["X_optm"].cpu().numpy()
    if args.gather == "peer":
        assert mpc.gather_error() == 0, "a gather wait timed out"
    # untimed check of the exchange: every rank's block of this rank's gathered set equals that rank's own solution
    g = unpack_flat_slab(sharded.gathered(last).cpu().numpy(), max(world, 1), args.batch, N, per=sharded.per)
    lo = rank * args.batch
    assert np.array_equal(g["X_optm"][lo:lo + args.batch], X_own)
    assert np.array_equal(g["status"][lo:lo + args.batch], status)
    if world > 1:
        chk = torch.from_numpy(np.array([float(X_own.sum()), float(status.sum())])).to(dev)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        for r in range(world):
            blk = g["X_optm"][r * args.batch:(r + 1) * args.batch]
            assert float(blk.sum()) == float(allc[r][0]), f"rank {rank}: block of rank {r} differs from what rank {r} solved"
            assert float(g["status"][r * args.batch:(r + 1) * args.batch].sum()) == float(allc[r][1])
    solved = int(((status == 0) | (status == 5)).sum())
    total_instances = args.batch * max(world, 1)
    value = total_instances * args.steps / (dev_ms * 1e-3)

    # ---- e2e: host (pinned) buffers through the C-ABI host path, wall clock around the synchronous call.  N > 1: the call
    # is lmpc_solve_gather_batch -- H2D, kernels with the fused exchange, wait for every peer, D2H of the gathered
    # trajectories of ALL ranks plus this rank's multipliers / safe-set columns.
    mpc.set_stream(None)
    h_in = mpc.alloc_host_inputs(data, pinned=True)     # one pinned arena in struct order: a single H2D copy per step
    h_out = mpc.alloc_host_outputs(args.batch, pinned=True)
    fused = world > 1 and args.gather == "peer"
    g_host = torch.empty(world * sharded.per, dtype=torch.float64).pin_memory().numpy() if fused else None

    def e2e_call(k, all_to_host=False):
        """N > 1: H2D of this rank's inputs, kernels with the fused exchange, wait for every peer (the gathered set is then
        valid in this rank's HBM), D2H of this rank's own outputs.  all_to_host: D2H of the trajectories of ALL ranks instead."""
        if fused:
            mpc.solve_gather(h_in, h_out, k % 2, wait=True, gathered_host=g_host if all_to_host else None)
        else:
            mpc.solve(h_in, h_out)    # N = 1, or the NCCL A/B modes (their collective is device-path only)

    def e2e_leg(all_to_host):
        for k in range(3):
            e2e_call(k, all_to_host)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_call(k + 3, all_to_host)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    e2e_s = e2e_leg(False)
    h2d = int(sum(v.nbytes for v in h_in.values()))
    d2h = int(sum(v.nbytes for v in h_out.values()))
    assert np.array_equal(h_out["status"], status0) and np.array_equal(h_out["X_optm"], X0)
    e2e_all = None
    if fused:
        assert mpc.gather_error() == 0
        s_all = e2e_leg(True)
        ge = unpack_flat_slab(g_host, world, args.batch, N, per=sharded.per)
        assert np.array_equal(ge["status"][lo:lo + args.batch], status0) and np.array_equal(ge["X_optm"][lo:lo + args.batch], X0)
        assert mpc.gather_error() == 0
        e2e_all = {"value": total_instances * args.steps / s_all, "unit": UNIT,
                   "d2h_bytes_per_step": int(g_host.nbytes + sum(h_out[k].nbytes for k in ("convex_combi_optm", "ss_x", "ss_j", "iters"))),
                   "note": "variant: every rank copies the trajectories of ALL ranks to its host (world x slab) instead of its own shard"}
    e2e_value = total_instances * args.steps / e2e_s

    # ---- roofline of the dominant kernel (lmpc_qp_kernel), live CUDA-event duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    fp64 = None
    if rank == 0:
        summary = load_summary()
        # fp64-pipe view: counted fp64 flops of the profiled launch (same seed and batch as this run,
        # profiles/qp_kernel_summary.json) over the live kernel time, against a DFMA rate measured on this device now
        try:
            dfma_peak = mpc.measure_fp64_peak()
            flops = summary.get("fp64_flops_per_launch") if summary.get("batch") == args.batch else None
            fp64 = {"peak_tflops": dfma_peak, "peak_source": "lmpc_measure_fp64_peak (DFMA chains, this device, this run)",
                    "flops_per_launch": flops,
                    "kernel_ms_of_the_profiled_batch": qp_ms_batch0,
                    "achieved_tflops": (flops / (qp_ms_batch0 * 1e-3) / 1e12) if flops and qp_ms_batch0 > 0 else None}
            fp64["frac"] = (fp64["achieved_tflops"] / dfma_peak) if fp64["achieved_tflops"] and dfma_peak > 0 else None
        except Exception as ex:   # measurement aid only
            fp64 = {"error": str(ex)}
    mpc.close()

    # ---- BASELINE configs[2..4] at their per-GPU share (every rank takes part: the lines are max-over-ranks too)
    configs = None
    if not args.no_configs:
        configs = {}
        for name in ("config3_iac_tracking_N40", "config4_50lap_regression", "config5_montecarlo_65536"):
            try:
                configs[name] = extra_config_line(pkg, name, dev, local_rank, stream, flush, dist, world, rank, peak)
            except Exception as ex:    # a failing side line must not take the headline with it
                configs[name] = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    alg_bytes = b_alg(N, K) * args.batch
    achieved = alg_bytes / (qp_ms * 1e-3) / 1e9 if qp_ms > 0 else 0.0
    roofline = {"kernel": "lmpc_qp_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": summary.get("dram_bytes_per_launch"),
                "traffic_source": "ncu --set full capture of the same launch (profiles/qp_kernel_summary.json), not re-measured in this run",
                "peak_source": peak_src,
                "algorithmic_bytes_per_step": b_alg(N, K), "kernel_ms": qp_ms,
                "kernel_ms_per_rank": per_rank_qp, "kernel_ms_min_max": [min(per_rank_qp), max(per_rank_qp)],
                "step_ms_per_rank": per_rank_step,
                "kernel_share_of_step": (ms_qp / max(ms_lin + ms_ss + ms_qp, 1e-12)),
                "other_kernels_ms": {"lmpc_linearise_kernel": ms_lin / max(nrec, 1), "lmpc_ss_query_kernel": ms_ss / max(nrec, 1)},
                "fp64": fp64,
                "note": "the path is fp64 latency/issue bound (about 300 flop/B): the HBM fraction is reported as the contract asks, "
                        "the fp64 object gives the pipe view (DESIGN.md)"}

    cpu = None
    if world == 1:
        res = cpu_arm(pkg, steps=3, warmup=1, sample_instances=4096)
        cpu = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
               "note": "CPU restatement -- reference stack (CasADi/OSQP) probed at run time and absent", "reference_stack_probe": ref_stack}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, 1), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_desc,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes_exchange": bool(fused) or world == 1,
                    "what": "per rank and step: H2D of the rank's inputs (pinned), the three kernels with the fused exchange, wait for every peer's results, D2H of the rank's own outputs; wall clock, max over ranks",
                    "all_ranks_to_every_host": e2e_all},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "solved_fraction": solved / args.batch, "ipm_iters_mean": float(iters.mean()), "ipm_iters_max": int(iters.max()),
            "configs": configs}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
