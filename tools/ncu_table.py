"""One line per captured launch of an .ncu-rep (kernel, time, DRAM bytes, pipe utilisations, shared-memory wavefronts).
Usage: python tools/ncu_table.py rep.ncu-rep [more.ncu-rep ...]   (runs here: ncu -i needs no GPU)"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tc%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_cf"),
        ("sm__cycles_elapsed.avg", "cycles"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block")]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-kernel-base", "demangled"],
                         capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units, data = r[0], r[1], r[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for d in data:
        name = re.sub(r"\(.*", "", d[ix["Kernel Name"]]).replace("oar::", "").replace("void ", "")
        rec = {"name": name}
        for c, short in COLS:
            if c not in ix:
                continue
            try:
                v = float(d[ix[c]].replace(",", ""))
            except ValueError:
                continue
            u = units[ix[c]]
            if short in ("rdMB", "wrMB"):
                v *= TO_MB.get(u, 1.0)
            if short == "us":
                v *= TO_US.get(u, 1.0)
            rec[short] = v
        yield rec


def main():
    for rep in sys.argv[1:]:
        print(f"# {rep}")
        print(f"{'kernel':44s} " + " ".join(f"{s:>8s}" for _, s in COLS))
        for rec in rows(rep):
            print(f"{rec['name'][:44]:44s} " + " ".join(
                (f"{rec[s]:8.1f}" if s in rec and rec[s] < 1e7 else (f"{rec[s]:8.2e}" if s in rec else f"{'-':>8s}"))
                for _, s in COLS))


if __name__ == "__main__":
    main()
