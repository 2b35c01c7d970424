"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept under profiles/.

usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rN_launch_summary.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ix = {k: j for j, k in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[ix['Metric Value']])
        u = r[ix['Metric Unit']]
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        name = r[ix['Kernel Name']]
        name = name.split('(')[0] if 'gemm_split' not in name else name.split('(CUtensorMap')[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(t for _, t in agg.values())
    n = sum(c for c, _ in agg.values())
    print("Total: %.1f ms of kernel time over %d launches.\n" % (tot / 1000, n))
    print("| launches | time (us) | share | kernel |\n|---:|---:|---:|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print("| %d | %.0f | %.1f%% | `%s` |" % (c, t, 100 * t / tot, k[:110]))
    gemm = sum(t for k, (c, t) in agg.items() if 'gemm_split' in k)
    ours = sum(t for k, (c, t) in agg.items() if 'nefii' in k)
    print("\nShare of the tcgen05 layer GEMM (all template variants): **%.1f%%** of kernel time; all nefii kernels %.1f%%." % (
        100 * gemm / tot, 100 * ours / tot))


if __name__ == "__main__":
    main(sys.argv[1])
