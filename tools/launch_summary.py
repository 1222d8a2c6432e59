"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for one
training step (delimited by the stft kernel).  usage: python tools/launch_summary.py list.csv [--md]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def one_step(rows):
    idx = [i for i, r in enumerate(rows) if "stft_mag" in r["Kernel Name"] or "stft_frames" in r["Kernel Name"]]
    return rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows       # the last complete step (a CUDA-graph replay in bench.py)


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "").replace("adp::", "")


def main():
    rows = [r for r in one_step(load(sys.argv[1])) if "spin_kernel" not in r["Kernel Name"]]   # (bench.py's queue-ahead spin)
    md = "--md" in sys.argv
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        t = float(r["Metric Value"]) / 1e3
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print("launches per step: %d, summed kernel time %.2f ms\n" % (len(rows), tot / 1e3))
    if md:
        print("| kernel | launches | us | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print("| `%s` | %d | %.1f | %.1f %% |" % (n[:70], c, t, 100 * t / tot))
        else:
            print("%-62s n=%3d %9.1f us %5.1f%%" % (n[:62], c, t, 100 * t / tot))
    if "--tc" in sys.argv:
        for r in rows:
            if "tc_" in r["Kernel Name"]:
                print("%-28s grid=%-18s %8.1f us" % (short(r["Kernel Name"]), r["Grid Size"], float(r["Metric Value"]) / 1e3))


if __name__ == "__main__":
    main()
