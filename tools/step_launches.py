"""One training step's kernel launches, in order, from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/step_launches.py launches.csv [other.csv]   (two files: side by side, for A/B of two builds)"""
import csv, re, sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"adp::<unnamed>::|adp::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*$", "", name)
        if "spin_kernel" in name:
            continue
        rows.append((name[:52], row["Grid Size"].replace(" ", ""), float(row["Metric Value"]) / 1000.0))
    starts = [i for i, r in enumerate(rows) if r[0].startswith("stft_frames_kernel")]
    if len(starts) >= 2:
        rows = rows[starts[-2]:starts[-1]]
    return rows


def main():
    a = load(sys.argv[1])
    b = load(sys.argv[2]) if len(sys.argv) > 2 else None
    if b is None:
        for n, g, t in a:
            print("%-54s %-14s %8.1f" % (n, g, t))
        print("total %.1f us over %d launches" % (sum(r[2] for r in a), len(a)))
        return
    print("A: %d launches %.1f us   B: %d launches %.1f us" % (len(a), sum(r[2] for r in a), len(b), sum(r[2] for r in b)))
    i = j = 0
    while i < len(a) or j < len(b):
        ra = a[i] if i < len(a) else None
        rb = b[j] if j < len(b) else None
        fam = lambda r: r[0].split("<")[0] if r else None
        if ra and rb and fam(ra) == fam(rb):
            print("%-46s %-12s %7.1f | %-46s %7.1f %+7.1f" % (ra[0][:46], ra[1], ra[2], rb[0][:46], rb[2], rb[2] - ra[2]))
            i += 1; j += 1
        elif rb and (not ra or fam(rb) not in [fam(x) for x in a[i:i + 6]]):
            print("%-46s %-12s %7s | %-46s %7.1f" % ("", "", "", rb[0][:46], rb[2]))
            j += 1
        else:
            print("%-46s %-12s %7.1f |" % (ra[0][:46], ra[1], ra[2]))
            i += 1


if __name__ == "__main__":
    main()
