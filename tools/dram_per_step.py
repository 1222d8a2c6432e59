"""Turn an ncu CSV of (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) per tc_* launch of ONE
training step into profiles/<name>.json (what bench.py reports as roofline.traffic).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:tc_ -c 800 --csv --log-file gpurun_out/tc_dram.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph
    python tools/dram_per_step.py gpurun_out/tc_dram.csv profiles/r2_tc_dram_per_step.json

The capture may hold several steps: the LAST complete one (from an STFT GEMM launch to the launch before the next) is used.
Template arguments: tc_igemm_persist_kernel<BLOCK_N, EG, ATT, HALO, ALT> (round 1 captures: <BLOCK_N, CLUSTER, EG, ATT, HALO>,
pass --r1 for those).
"""
import csv
import io
import json
import sys


def main(src, dst, r1=False):
    txt = open(src).read().splitlines()
    start = [k for k, line in enumerate(txt) if line.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[start:]))))
    by_id = {}
    for r in rows:
        e = by_id.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            e["read" if "read" in r["Metric Name"] else "write"] = v * scale
        else:
            e["us"] = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}[unit]
    launches = list(by_id.values())
    eg_pos = 2 if r1 else 1
    stft_at = [i for i, e in enumerate(launches)
               if (lambda m: bool(m) and m.group(1).split(",")[eg_pos].strip().endswith("2"))(
                   __import__("re").search(r"tc_igemm_persist_kernel<([^>]*)>", e["kernel"]))]
    if len(stft_at) >= 2:          # several steps captured: keep the last complete one
        launches = launches[stft_at[-2]:stft_at[-1]]

    def family(e):
        k = e["kernel"]
        if "tc_wgrad" in k:
            return "conv"
        if "tc_igemm_persist_kernel" in k:
            return "conv"
        return "other"

    # the conv families of bench.py's roofline = gather + parity + wgrad over E2-E8 / D8-D2; the thin first/last layers
    # (pointwise GEMMs: <16>, gemm_tn, the two 1M-pixel <64> pointwise launches) and the STFT GEMM are listed but not averaged
    per = []
    for e in launches:
        name = e["kernel"].split("(")[0].replace("void adp::<unnamed>::", "")
        per.append({"kernel": name, "grid": e["grid"], "read_MB": round(e.get("read", 0) / 1e6, 1),
                    "write_MB": round(e.get("write", 0) / 1e6, 1), "us": round(e.get("us", 0), 1)})
    import re

    def targs(k):      # template arguments of tc_igemm_persist_kernel<BLOCK_N, CLUSTER, EG, ATT, HALO>
        m = re.search(r"tc_igemm_persist_kernel<([^>]*)>", k)
        return [int(re.sub(r"\([a-z]+\)", "", a).strip()) for a in m.group(1).split(",")] if m else None

    def is_stft(e):
        a = targs(e["kernel"])
        return bool(a) and len(a) >= 3 and a[eg_pos] == 2

    def is_thin(e):
        a = targs(e["kernel"])
        return "gemm_tn" in e["kernel"] or is_stft(e) or (bool(a) and a[0] == 16)
    thin = [i for i, e in enumerate(per) if is_thin(e)]
    # the E1 pointwise GEMM is the launch right after the STFT GEMM, the last-layer dgrad the one right after gemm_tn<64>
    # (round 1 only: since round 2 the thin layers run in adp_thin_tc.cu's own kernels, which the tc_ filter does not capture)
    for i, e in enumerate(per[:-1] if r1 else []):
        if is_stft(e) or "gemm_tn_kernel<64>" in e["kernel"] or "gemm_tn_kernel<(int)64>" in e["kernel"]:
            thin.append(i + 1)
    conv = [i for i in range(len(per)) if i not in thin]
    tot = sum((launches[i].get("read", 0) + launches[i].get("write", 0)) for i in conv)
    out = {"what": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the tcgen05 launches of ONE training step, "
                   "B=64 bf16, 1xB200; averaged over the conv-family launches (everything but the STFT GEMM, gemm_tn and "
                   "the <16> head GEMM)",
           "launches": len(conv), "dram_bytes_per_step": tot, "dram_bytes_per_launch": tot / max(len(conv), 1),
           "ncu_time_us_per_step": sum(per[i]["us"] for i in conv), "per_launch": per}
    json.dump(out, open(dst, "w"), indent=1)
    print(dst, "launches", len(conv), "MB/launch %.1f" % (out["dram_bytes_per_launch"] / 1e6), "us", out["ncu_time_us_per_step"])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], "--r1" in sys.argv)
