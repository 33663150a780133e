"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed).
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/<name>_launches.txt
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep > profiles/<name>_metrics.txt"""
import collections
import csv
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:70]
        v = float(r[mv].replace(",", ""))
        v = v / 1e6 if r[mu] == "ns" else (v / 1e3 if r[mu] == "us" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: gpu__time_duration.sum per kernel (ms; cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} n={n:4d} total={t:9.3f} ms  {100 * t / tot:5.1f}%")
    print(f"total {tot:.3f} ms")


def raw(path):
    if path.endswith(".csv"):          # a raw page exported on the GPU box (the reports themselves are too big to travel)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    idx = {k: h.index(k) for k in KEY if k in h}
    kn = h.index("Kernel Name")
    print(f"# {path}: ncu --set full, selected metrics per captured launch")
    for r in rows[2:]:
        print("----", r[kn][:90])
        for k in KEY:
            if k in idx:
                print(f"   {k:82s} {r[idx[k]]:>18s} {rows[1][idx[k]]}")


def traffic(path, key, out_json):
    """Append dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch in an `ncu --set full` report to
    the JSON table bench.py reads (profiles/r02_traffic.json): key = "<kernel>|M..,D..,R..,N..".""" 
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, r = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(name)
        tot += float(r[i].replace(",", "")) * scale[units[i]]
    tbl = json.load(open(out_json)) if os.path.exists(out_json) else {}
    tbl[key] = tot
    json.dump(tbl, open(out_json, "w"), indent=1, sort_keys=True)
    print(key, tot)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
