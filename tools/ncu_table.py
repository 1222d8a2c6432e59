"""Tabulate the key metrics of an `ncu --set full` report: python tools/ncu_table.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("launch__grid_size", "grid"),
    ("dram__bytes_read.sum", "rd"),
    ("dram__bytes_write.sum", "wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2B"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe%"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16_ops%"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tmem%"),
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
    print("kernel," + ",".join("%s[%s]" % (n, units[idx[m]]) for m, n in cols))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void adp::<unnamed>::", "").replace("void <unnamed>::", "")
        print(name[:28] + "," + ",".join(r[idx[m]] for m, n in cols))


if __name__ == "__main__":
    main()
