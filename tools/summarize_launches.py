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


def phases(path):
    """Groups the SDF evaluations of a fixed-schedule launch list (encode_kernel + 8 layer GEMMs each) by the tracer kernel that
    asked for them; an evaluation counts as "with work" when its kernels ran for more than 50 us in total."""
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    ix = {k: j for j, k in enumerate(rows[h])}
    seq = []
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[ix['Metric Value']])
        u = r[ix['Metric Unit']]
        seq.append((r[ix['Kernel Name']], v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v))
    kinds = (('march_round_kernel', 'march rounds'), ('bisect_kernel', 'bisection rounds'), ('sample_emit_kernel', '100-sample scans (sampler, min-SDF)'))
    agg = collections.defaultdict(lambda: [0, 0, 0.0])
    other = 0.0
    i = 0
    last_tracer = None
    while i < len(seq):
        name, t = seq[i]
        if 'encode_kernel' in name:
            j = i + 1
            tot = t
            while j < len(seq) and 'gemm_split' in seq[j][0]:
                tot += seq[j][1]
                j += 1
            key = last_tracer or 'stand-alone evaluations (sdf_output, features + normals on hits)'
            agg[key][0] += 1
            if tot > 50.0:
                agg[key][1] += 1
            agg[key][2] += tot
            last_tracer = None
            i = j
            continue
        for pat, label in kinds:
            if pat in name:
                last_tracer = label
        other += t
        i += 1
    print("| phase | evaluations launched | with work | time (ms) |\n|---|---:|---:|---:|")
    for k, (n, w, t) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print("| %s | %d | %d | %.1f |" % (k, n, w, t / 1000))
    print("| everything that is not an SDF evaluation | | | %.1f |" % (other / 1000))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--phases":
        phases(sys.argv[2])
    else:
        main(sys.argv[1])
