"""Condense ncu outputs from gpurun_out/ into small text summaries for profiles/ (tracked).
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep   > profiles/rNN_conv_full.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum [", "dram__bytes_write.sum [",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread [",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum [",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]


def launches(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)")
    print(f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} n={n:4d} total_us={t:10.1f} avg_us={t / n:8.1f} share={100 * t / tot:5.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    hdr, units = next(rd), next(rd)
    idx = [i for i, h in enumerate(hdr) if any((h + " [").startswith(k) or h.startswith(k) for k in KEYS)]
    it = hdr.index("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None
    ip = next((i for i, h in enumerate(hdr) if h.startswith("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")), None)
    idr = next((i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum") and units[i].endswith("byte")), None)
    idw = next((i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum") and units[i].endswith("byte")), None)
    tw = tt = dram = 0.0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for n, row in enumerate(rd):
        print(f"--- launch {n}")
        for i in idx:
            print(f"  {hdr[i]} [{units[i]}] = {row[i]}")
        try:
            t = float(row[it].replace(",", ""))
            tt += t
            tw += t * float(row[ip].replace(",", ""))
            dram += float(row[idr].replace(",", "")) * scale.get(units[idr], 1.0) + float(row[idw].replace(",", "")) * scale.get(units[idw], 1.0)
        except Exception:
            pass
    if tt > 0:
        print(f"=== time-weighted sm__pipe_tensor_cycles_active over these launches: {tw / tt:.1f} %  (total {tt:.4f} {units[it]}, dram read+write {dram / 1e9:.3f} GB)")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
