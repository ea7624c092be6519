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
        orc.step_batch({k: v[:max(cores, 64)] for k, v in data.items()}, impl="port", nthreads=cores)
    t0 = time.perf_counter()
    nfail = 0
    for _ in range(steps):
        r = orc.step_batch(data, impl="port", nthreads=cores)
        nfail += int(r["nfail"])
    dt = time.perf_counter() - t0
    return dict(value=steps * sample_instances / dt, seconds=dt, cores=cores, failed=nfail,
                sample=f"{steps} x {sample_instances} instances of the same workload (oracle port, {cores} threads)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="instances per GPU (default: BASELINE config 2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import racing_lmpc_ros2_b200 as pkg

    config_desc = {"workload": f"BASELINE configs[1]: BARC LMPC, N={N_HORIZON}, 6-state Frenet bicycle, K=96 safe-set columns "
                               f"(3 recorded laps), {args.batch} random initial states per GPU",
                   "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1), "N": N_HORIZON, "K": 96,
                   "parallelism": f"instances sharded over {max(world, 1)} GPU(s), safe set replicated, one all-gather of trajectories per step (overlapped with the next step's solve, waited inside its timed window)",
                   "l2": "256 MiB scratch written between timed steps (untimed) to flush the 126 MB L2", "tol": 1e-7, "polish": "active-set (augmented-Lagrangian) polish after the interior point"}

    # ------------------------------------------------------------------ CPU ("reference") arm
    if args.impl == "reference":
        if rank != 0:
            return
        sample = 4096
        res = cpu_arm(pkg, steps=max(1, min(args.steps, 20)), warmup=1, sample_instances=sample)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.batch * max(world, 1) / res["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_desc,
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
                                 "note": "CPU restatement -- reference stack (CasADi/OSQP) unavailable in this environment"},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solve path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    veh, cfg, track, laps, data = workload(pkg, 0xB200 + 2 + 7919 * rank, args.batch)
    mpc = BatchedRacingMPC(veh, cfg, max_batch=args.batch, device=local_rank)
    # safe set: rank 0 owns the laps, every rank receives them (replicated), then ingests locally
    for l in laps:
        x = torch.from_numpy(np.ascontiguousarray(l["x"])).to(dev)
        if world > 1:
            dist.broadcast(x, src=0)
        mpc.add_lap(x.cpu().numpy(), l["u"], l["k"], l["t"], track["length"])

    stream = torch.cuda.Stream(device=dev)
    mpc.set_stream(stream)
    d_in = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    # two output sets: while the trajectories of step k are gathered (NCCL, its own stream), step k + 1 solves into the other
    d_outs = [mpc.alloc_device_outputs(args.batch, dev) for _ in range(2 if world > 1 else 1)]
    d_out = d_outs[0]
    N, K = cfg["N"], cfg["num_ss_pts"]
    # X, U, dU, cost, status of the rank live in one allocation (d_out["slab"]): the gather needs no packing kernel
    gathered = [torch.empty(world * d_out["slab"].numel(), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    pending = [None]

    def step(k):
        """Solve, then the one collective of the path (all-gather of the converged trajectories).  The gather of step k
        overlaps the solve of step k + 1 and is waited for INSIDE that step's timed window (the last one in a window of
        its own, drain()), so every gather is inside the timed region."""
        o = d_outs[k % len(d_outs)]
        mpc.solve(d_in, o)
        if world > 1:
            if pending[0] is not None:
                pending[0].wait()
            pending[0] = dist.all_gather_into_tensor(gathered[k % 2], o["slab"], async_op=True)

    def drain():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            step(k)
        drain()
    stream.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    mpc.set_timing(True)
    launches0 = mpc.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev_tail = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.fill_(float(k))            # untimed L2 flush
            ev[k][0].record(stream)
            step(k)
            ev[k][1].record(stream)
        ev_tail[0].record(stream)            # the gather of the last step, timed on its own
        drain()
        ev_tail[1].record(stream)
    stream.synchronize()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    launches = mpc.launch_count - launches0
    (ms_lin, ms_ss, ms_qp), nrec = mpc.kernel_ms()
    mpc.set_timing(False)
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) + ev_tail[0].elapsed_time(ev_tail[1])
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    last = (args.steps - 1) % len(d_outs)
    d_out = d_outs[last]
    status = d_out["status"].cpu().numpy()
    iters = d_out["iters"].cpu().numpy()
    if world > 1:   # untimed check of the collective: this rank's block of the gathered buffer is its own solution
        from racing_lmpc_ros2_b200.distributed import unpack_flat_slab
        g = unpack_flat_slab(gathered[(args.steps - 1) % 2].cpu().numpy(), world, args.batch, N)
        lo = rank * args.batch
        assert np.array_equal(g["X_optm"][lo:lo + args.batch], d_out["X_optm"].cpu().numpy())
        assert np.array_equal(g["status"][lo:lo + args.batch], status)
    solved = int((status == 0).sum())
    total_instances = args.batch * max(world, 1)
    value = total_instances * args.steps / (dev_ms * 1e-3)

    # ---- e2e: host (pinned) buffers through the C-ABI host path, wall clock around the synchronous call
    mpc.set_stream(None)
    h_in = mpc.alloc_host_inputs(data, pinned=True)     # one pinned arena in struct order: a single H2D copy per step
    h_out = mpc.alloc_host_outputs(args.batch, pinned=True)
    for _ in range(3):
        mpc.solve(h_in, h_out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mpc.solve(h_in, h_out)      # H2D of every input, 3 kernels, D2H of every output (safe-set columns behind the QP kernel), stream sync
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = int(sum(v.nbytes for v in h_in.values()))
    d2h = int(sum(v.nbytes for v in h_out.values()))
    e2e_value = total_instances * args.steps / e2e_s
    assert np.array_equal(h_out["status"], status)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (lmpc_qp_kernel), live CUDA-event duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    qp_ms = ms_qp / max(nrec, 1)
    alg_bytes = b_alg(N, K) * args.batch
    achieved = alg_bytes / (qp_ms * 1e-3) / 1e9 if qp_ms > 0 else 0.0
    traffic = None
    summary = {}
    summ = os.path.join(ROOT, "profiles", "qp_kernel_summary.json")
    if os.path.exists(summ):
        try:
            summary = json.load(open(summ))
            traffic = summary.get("dram_bytes_per_launch")
        except Exception:
            summary = {}
    # fp64-pipe view: counted fp64 flops of the profiled launch (same seed and batch as this run,
    # profiles/qp_kernel_summary.json) over the live kernel time, against a DFMA rate measured on this device now
    fp64 = None
    try:
        dfma_peak = mpc.measure_fp64_peak()
        flops = summary.get("fp64_flops_per_launch") if summary.get("batch") == args.batch else None
        fp64 = {"peak_tflops": dfma_peak, "peak_source": "lmpc_measure_fp64_peak (DFMA chains, this device, this run)",
                "flops_per_launch": flops,
                "achieved_tflops": (flops / (qp_ms * 1e-3) / 1e12) if flops and qp_ms > 0 else None}
        fp64["frac"] = (fp64["achieved_tflops"] / dfma_peak) if fp64["achieved_tflops"] and dfma_peak > 0 else None
    except Exception as ex:   # measurement aid only
        fp64 = {"error": str(ex)}
    roofline = {"kernel": "lmpc_qp_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_step": b_alg(N, K), "kernel_ms": qp_ms,
                "kernel_share_of_step": (ms_qp / max(ms_lin + ms_ss + ms_qp, 1e-12)),
                "other_kernels_ms": {"lmpc_linearise_kernel": ms_lin / max(nrec, 1), "lmpc_ss_query_kernel": ms_ss / max(nrec, 1)},
                "fp64": fp64,
                "note": "the path is fp64 latency/issue bound (about 300 flop/B): the HBM fraction is reported as the contract asks, "
                        "the fp64 object gives the pipe view (DESIGN.md)"}

    cpu = None
    if world == 1:
        res = cpu_arm(pkg, steps=3, warmup=1, sample_instances=4096)
        cpu = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"],
               "note": "CPU restatement -- reference stack (CasADi/OSQP) unavailable in this environment"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, 1), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_desc,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "solved_fraction": solved / args.batch, "ipm_iters_mean": float(iters.mean()), "ipm_iters_max": int(iters.max())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
