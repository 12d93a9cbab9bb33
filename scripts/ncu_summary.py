"""Summarise an ncu --set full report (read here, no GPU needed) into profiles/:

    python scripts/ncu_summary.py gpurun_out/prof_sweep.ncu-rep profiles/r1_v4_ncu_full_summary.csv [factors]

Writes the per-launch metric table and, for the k_sweep launch, profiles/traffic_k_sweep.json
(dram__bytes_read.sum + dram__bytes_write.sum of that launch), which bench.py reports as roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    factors = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, launches = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    names = ["Kernel Name", "Block Size", "Grid Size"] + [m for m in METRICS if m in col]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        for n in names:
            w.writerow([n, units[col[n]]] + [r[col[n]] for r in launches])
    for r in launches:
        if "k_sweep" in r[col["Kernel Name"]]:
            def mb(name):
                v, u = float(r[col[name]]), units[col[name]]
                return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u]
            rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
            tr = {"kernel": r[col["Kernel Name"]], "factors": factors, "dram_bytes_per_launch": rd + wr,
                  "dram_bytes_read": rd, "dram_bytes_write": wr,
                  "gpu_time_us_under_ncu": float(r[col["gpu__time_duration.sum"]]),
                  "source": f"{out} (ncu --set full --clock-control none, one non-relinearising launch)"}
            with open(os.path.join(os.path.dirname(out), "traffic_k_sweep.json"), "w") as f:
                json.dump(tr, f, indent=1)
            print(json.dumps(tr))
            break


if __name__ == "__main__":
    main()
