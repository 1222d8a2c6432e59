"""Tabulate the key metrics of an `ncu --set full` report: python tools/ncu_table.py report.ncu-rep

Tensor-core activity of tcgen05 kernels: `tensor_active%` = TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg
/ 4 sub-partitions / sm__cycles_elapsed.avg -- the cycles the bf16 MMA sub-pipes were busy (UTCHMMA is counted there; it equals
algorithmic FLOP / (8192 FLOP/clk/SM) on every launch checked).  `mem_tensor%` (sm__mem_tensor_cycles_active) tracks it within a
point or two.  The `...pipe_tensor_cycles_active_realtime.pct` and `sm__ops_path_tensor_op_hmma_*` columns of ncu 2025.2 do NOT
count UTCHMMA correctly (they read 0-30 % / 0 on kernels that run at 65 % of peak) and are not printed."""
import csv
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("launch__grid_size", "grid"),
    ("dram__bytes_read.sum", "rd"),
    ("dram__bytes_write.sum", "wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2B"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "mem_tensor%"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_inst%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    tens = [h for h in hdr if "tensor" in h and "pct" in h]
    if "--list-tensor" in sys.argv:
        print("\n".join(tens))
    cols = [(m, n) for m, n in WANT if m in idx]
    hm, cy = "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg"
    print("kernel,tensor_active[%]," + ",".join("%s[%s]" % (n, units[idx[m]]) for m, n in cols))
    for r in rows[2:]:
        name = re.sub(r"^void |adp::|<unnamed>::|unnamed>::|\(anonymous namespace\)::", "", r[idx["Kernel Name"]].split("(")[0])
        ta = ""
        try:
            ta = "%.1f" % (100.0 * float(r[idx[hm]]) / 4.0 / float(r[idx[cy]]))
        except Exception:
            pass
        print(name[:44] + "," + ta + "," + ",".join(r[idx[m]] for m, n in cols))


if __name__ == "__main__":
    main()
