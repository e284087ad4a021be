"""Summarises ncu outputs into small text files for profiles/ (run here, no GPU needed).

  python profiles/summarize.py launches gpurun_out/x_launches.csv  > profiles/x_launches.txt
  python profiles/summarize.py full gpurun_out/x.ncu-rep            > profiles/x_full.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name).replace("oar::", "")
        rows.append((name, v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1000:.2f} ms of kernel time (ncu: cold cache, serialised)")
    print(f"{'kernel':48s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:48]:48s} {c:8d} {us / 1000:10.3f} {us / c:9.1f} {us / total:7.3f}")


KEYS = ("gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct|"
        "sm__throughput.avg.pct|sm__pipe_tensor.*cycles_active.*pct|sm__warps_active.avg.pct|"
        "launch__registers_per_thread|launch__grid_size|launch__block_size|launch__shared_mem_per_block|"
        "launch__occupancy_limit|sm__inst_executed_pipe_tensor|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate.pct|"
        "smsp__cycles_active.avg|launch__waves_per_multiprocessor|sm__inst_executed_pipe_uniform")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rdr = list(csv.reader(io.StringIO(out)))
    hdr, units = rdr[0], rdr[1]
    pat = re.compile(KEYS)
    cols = [i for i, h in enumerate(hdr) if pat.search(h)]
    name_i = hdr.index("Kernel Name")
    print(f"# {path}")
    for row in rdr[2:]:
        print(f"## {row[name_i][:100]}")
        for i in cols:
            print(f"  {hdr[i]:60s} {row[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
