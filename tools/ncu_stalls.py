"""Warp-stall breakdown (stall cycles per issued instruction, by reason) of every launch in an .ncu-rep.
Usage: python tools/ncu_stalls.py rep.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
    if len(sys.argv) > 2:
        cmd += ["--kernel-name", "regex:" + sys.argv[2]]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h, rows = r[0], r[2:]
    ix = {k: i for i, k in enumerate(h)}
    keys = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
            and "not_issued" not in k]
    for d in rows:
        name = re.sub(r"\(.*", "", d[ix["Kernel Name"]])
        vals = []
        for k in keys:
            try:
                v = float(d[ix[k]])
            except ValueError:
                continue
            if v >= 0.1:
                vals.append((v, k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        vals.sort(reverse=True)
        print(name[:50], d[ix["gpu__time_duration.sum"]], " ".join(f"{n}={v:.2f}" for v, n in vals))


if __name__ == "__main__":
    main()
