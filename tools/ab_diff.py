"""Diff two `tools/layerprof.py --out` files (library A vs library B built from different sources, selected with
OAR_B200_LIB): per (kernel, bytes-per-launch) row the time in A and B, the delta, and the totals.

Typical round-2 use, ONE gpurun call for an A/B of a kernel change:
    cp oar_ocr_b200/liboar_b200.so oar_ocr_b200/_build/a.so     # before the change (in-tree, so it travels to the box)
    ... edit, python -m oar_ocr_b200.build ...
    gpurun -- 'OAR_B200_LIB=/root/repo/oar_ocr_b200/_build/a.so python tools/layerprof.py --out gpurun_out/a.json;
               python tools/layerprof.py --out gpurun_out/b.json; python tools/ab_diff.py gpurun_out/a.json gpurun_out/b.json'
Rows are matched by (name, MB per launch) in order of appearance, so layers keep their identity across builds as long
as the change does not alter a layer's algorithmic bytes; unmatched rows are listed separately."""
import json
import sys


def load(path):
    rows = json.load(open(path))
    out = {}
    for r in rows:
        key = (r["name"], r["mb"])
        n = 0
        while (key, n) in out:
            n += 1
        out[(key, n)] = r
    return out


def main():
    if len(sys.argv) != 3:
        print(__doc__)
        return 2
    a, b = load(sys.argv[1]), load(sys.argv[2])
    ta = sum(r["ms"] for r in a.values())
    tb = sum(r["ms"] for r in b.values())
    print(f"{'kernel':22s} {'MB/launch':>10s} {'A ms':>8s} {'B ms':>8s} {'delta':>8s} {'B GB/s':>8s}")
    by_name = {}
    for k, ra in a.items():
        rb = b.get(k)
        if rb is None:
            continue
        d = rb["ms"] - ra["ms"]
        if abs(d) >= 0.005:
            print(f"{ra['name']:22s} {ra['mb']:10.2f} {ra['ms']:8.3f} {rb['ms']:8.3f} {d:+8.3f} {rb['gbs']:8.1f}")
        s = by_name.setdefault(ra["name"], [0.0, 0.0])
        s[0] += ra["ms"]
        s[1] += rb["ms"]
    only_a = [r for k, r in a.items() if k not in b]
    only_b = [r for k, r in b.items() if k not in a]
    for tag, rows in (("only in A", only_a), ("only in B", only_b)):
        for r in rows:
            print(f"{tag}: {r['name']:22s} n={r['n']} ms={r['ms']:.3f} MB/launch={r['mb']}")
    print("\nper kernel name (matched rows):")
    for name, (x, y) in sorted(by_name.items(), key=lambda kv: kv[1][0] - kv[1][1], reverse=True):
        if abs(y - x) >= 0.005:
            print(f"  {name:22s} {x:8.3f} -> {y:8.3f}  ({y - x:+.3f} ms)")
    print(f"\ntotal kernel ms: A {ta:.3f}  B {tb:.3f}  ({tb - ta:+.3f} ms, {100.0 * (tb - ta) / ta:+.2f} %)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
