"""Where a warp of the sweep kernel spends its time, from an ncu --set full report (read here, no GPU needed):

    python scripts/ncu_segments.py gpurun_out/r2_prof_sweep.ncu-rep > profiles/r2_k_sweep_tma_ncu_summary.txt

Prints the speed-of-light percentages and stall ratios of the first sweep-kernel launch, then the stall samples of its
hot loop grouped into code segments (cut at the exact divisions / barriers / stores)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head = rows[0]
launch = [r for r in rows[2:] if "k_sweep" in r[head.index("Kernel Name")]][0]
d = dict(zip(head, launch))
print("kernel:", d["Kernel Name"], "| block", d.get("Block Size"), "| grid", d.get("Grid Size"))
for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
          "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if k in d:
        print(f"  {k:70s} {d[k]}")
print("stall cycles per issued instruction:")
for k in head:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        print(f"  {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {float(d[k]):.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
ks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        ks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
k = [k for k in ks if "k_sweep" in k["name"]][0]
h = k["hdr"]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ix["# Samples"]]) for r in k["rows"])
hot = max(int(r[ix["Instructions Executed"]]) for r in k["rows"])
print(f"\nstall samples of the hot loop by code segment ({tot} samples in all; instructions executed >= {hot // 2} times):")
seg, cur = [], None
for r in k["rows"]:
    if int(r[ix["Instructions Executed"]]) < hot // 2:
        continue
    op = [t for t in r[ix["Source"]].split() if not t.startswith("@")][0].split(".")[0]
    if cur is None:
        cur = {"n": 0, "s": 0, "st": collections.Counter(), "ops": collections.Counter(), "start": r[ix["Address"]][-5:]}
    cur["n"] += 1
    cur["s"] += int(r[ix["# Samples"]])
    cur["ops"][op] += 1
    for s in stalls:
        cur["st"][s[6:]] += int(r[ix[s]])
    if op in ("FCHK", "LDGDEPBAR", "SYNCS", "WARPSYNC", "STG") and cur["n"] >= 25:
        seg.append(cur)
        cur = None
if cur:
    seg.append(cur)
for s in seg:
    top = ", ".join(f"{a}:{b}" for a, b in s["st"].most_common(4))
    ops = ", ".join(f"{a}:{b}" for a, b in s["ops"].most_common(4))
    print(f"  @{s['start']} {s['n']:4d} instr {100 * s['s'] / tot:5.1f} % of samples ({s['s'] / max(s['n'], 1):5.2f} per instr) | {top} | {ops}")
