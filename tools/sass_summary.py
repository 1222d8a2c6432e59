"""SASS summary of the shipped library: python tools/sass_summary.py [label] > profiles/<round>_sass_summary.txt
Counts the Blackwell tensor-core / TMA / TMEM mnemonics over `cuobjdump -sass` of libadp_b200.so and lists them per kernel
(the proof that the hot kernels issue tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, cp.async.bulk.tensor = UTMALDG)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "audio_depth_estimation_b200", "libadp_b200.so")
PATTERNS = [
    ("UTCHMMA", r"\bUTCHMMA\b"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
    ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("UTCBAR", r"\bUTCBAR\b"),
    ("UTCATOMSWS (tmem alloc)", r"\bUTCATOMSWS\b"), ("HMMA (legacy)", r"\bHMMA\b"), ("SYNCS (mbarrier)", r"\bSYNCS\b"),
    ("FENCE.VIEW.ASYNC (generic -> async proxy)", r"\bFENCE\.VIEW\.ASYNC\b"), ("REDG.E.ADD.F32 (split-K / dw)", r"\bREDG?\.E\.ADD\.F32"),
]


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "current build"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    total = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"^_ZN3adp\d*_?GLOBAL__N__[0-9a-f]+_\d+_", "", m.group(1))
            cur = re.sub(r"^_ZN3adp\d+", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        body = line.split("*/", 1)[-1]          # (skip the address comment)
        for name, pat in PATTERNS:
            if re.search(pat, body):
                total[name] += 1
                per[cur][name] += 1
    tc = [(k, c) for k, c in per.items() if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]]
    print("SASS summary of audio_depth_estimation_b200/libadp_b200.so (cuobjdump -sass, sm_100a) -- %s" % label)
    print("kernels in the library: %d; kernels that issue tcgen05 / TMA / TMEM instructions: %d\n" % (len(per), len(tc)))
    for name, _ in PATTERNS:
        print("%-45s %6d" % (name, total[name]))
    print("\n%-100s %8s %6s %8s %7s" % ("kernel (mangled, namespace stripped)", "UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"))
    for k, c in sorted(tc):
        print("%-100s %8d %6d %8d %7d" % (k[:100], c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UTCBAR"]))


if __name__ == "__main__":
    main()
