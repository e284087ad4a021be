"""Summarise an .ncu-rep: per-launch headline metrics, and (optionally) the stall profile of one launch by SASS range.
Usage: python tools/ncu_summary.py rep.ncu-rep [launch_index_for_source_page]"""
import csv
import io
import subprocess
import sys

WANT = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__warps_active.avg.per_cycle_active']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for w in WANT:
        if w in ix:
            print(f"{w} [{units[ix[w]]}]:", [d[ix[w]][:28] for d in data])


def source(rep, launch):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(launch), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][:2])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, d = set(), []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == 'Address' or r[0] in seen:
            continue
        seen.add(r[0])
        d.append(r)

    def F(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(F(r, '# Samples') for r in d)
    print('instructions', len(d), 'samples', tot)
    step = 100
    for a in range(0, len(d), step):
        seg = d[a:a + step]
        s = sum(F(r, '# Samples') for r in seg)
        if s < tot * 0.004:
            continue
        agg = sorted(((st, sum(F(r, st) for r in seg)) for st in stalls), key=lambda x: -x[1])[:3]
        ex = sum(F(r, 'Instructions Executed') for r in seg)
        print(f"{a:5d} samples={s:7.0f} exec={ex:10.0f} {seg[0][ix['Source']][:40]:40s}", [(k[6:], int(v)) for k, v in agg])
    top = sorted(d, key=lambda r: -F(r, '# Samples'))[:24]
    for r in top:
        st = sorted(((s, F(r, s)) for s in stalls), key=lambda x: -x[1])[:2]
        print(int(F(r, '# Samples')), d.index(r), r[ix['Source']][:70], [(k[6:], int(v)) for k, v in st], r[ix['Instructions Executed']])


if __name__ == '__main__':
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[1], int(sys.argv[2]))
