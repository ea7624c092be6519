"""Condense an `ncu --set full` report of lmpc_qp_kernel into profiles/<name>.json.

usage: python scripts/ncu_summary.py gpurun_out/qp_rXX.ncu-rep profiles/qp_kernel_summary.json [batch]

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps the handful of numbers DESIGN.md / bench.py cite:
duration, DRAM traffic per launch, occupancy limits, issue / fp64-pipe utilisation, counted fp64 flops, stall mix.
"""
import csv, io, json, subprocess, sys


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = rows[2:]
    v = launches[0]

    def col(name):
        i = hdr.index(name)
        x = float(v[i].replace(",", ""))
        u = units[i]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                 "Kbyte/block": 1e3, "byte/block": 1.0}.get(u, 1.0)
        return x * scale

    def opt(name):
        try:
            return col(name)
        except (ValueError, IndexError):
            return None

    cyc = col("smsp__cycles_elapsed.avg")
    per_cycle = {k: col(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") for k in ("dfma", "dadd", "dmul")}
    flops = (2 * per_cycle["dfma"] + per_cycle["dadd"] + per_cycle["dmul"]) * cyc
    stalls = {}
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            x = opt(h)
            if x and x > 0.02:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(x, 3)
    kname = v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "lmpc_qp_kernel"
    out = {
        "source": rep, "kernel": kname.split("(")[0], "batch": batch, "launches_in_report": len(launches),
        "note": "captured under ncu --set full --clock-control none (replayed, cold cache): use shares and counts, not the absolute time",
        "duration_ms_under_ncu": col("gpu__time_duration.sum") * 1e3,
        "dram_bytes_read": col("dram__bytes_read.sum"), "dram_bytes_write": col("dram__bytes_write.sum"),
        "dram_bytes_per_launch": col("dram__bytes_read.sum") + col("dram__bytes_write.sum"),
        "registers_per_thread": col("launch__registers_per_thread"),
        "shared_mem_per_block_bytes": col("launch__shared_mem_per_block"),
        "occupancy_limit_blocks_per_sm": {"registers": col("launch__occupancy_limit_registers"), "shared_mem": col("launch__occupancy_limit_shared_mem"),
                                          "warps": col("launch__occupancy_limit_warps")},
        "warps_active_pct_of_peak": col("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_cycles_active_pct_elapsed": col("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "fp64_thread_inst_per_cycle": per_cycle,
        "fp64_flops_per_launch": flops,
        "fp64_flops_per_instance": flops / batch,
        "warp_inst_executed": col("smsp__inst_executed.sum"),
        "warp_inst_per_instance": col("smsp__inst_executed.sum") / batch,
        "active_threads_per_warp_inst": opt("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "local_load_inst": opt("sass__inst_executed_local_loads"), "local_store_inst": opt("sass__inst_executed_local_stores"),
        "stall_cycles_per_issue": stalls,
    }
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
