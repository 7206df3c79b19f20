#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,...,dram__bytes_*.sum --csv` launch list of bench.py:
picks one complete forward (image staging kernel .. policy head), writes a per-launch breakdown CSV and the DRAM
traffic of the tensor-core convolution launches (bench.py's roofline.traffic reads the JSON).

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r1_forward_breakdown.csv profiles/conv_traffic.json
"""
import csv
import json
import sys


def main(src, out_csv, out_json):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    by = {}
    for r in rows:
        d = by.setdefault(int(r[0]), {"name": r[4], "grid": r[8]})
        d[r[12]] = float(r[14])
    ids = sorted(by)
    starts = [i for i in ids if "image_nchw" in by[i]["name"]]
    if len(starts) < 2:
        raise SystemExit("need a launch list that spans at least one complete forward")
    lo, hi = starts[-2], starts[-1]
    fwd = [by[i] for i in ids if lo <= i < hi]
    conv = [d for d in fwd if any(k in d["name"] for k in ("conv_tc_kernel", "conv3x3_flat_kernel", "stem_pool_kernel", "stem_tc_kernel"))]
    with open(out_csv, "w") as f:
        f.write("# one forward of bench.py (batch 256, eager launches under ncu: cold-cache, serialised - compare SHARES)\n")
        f.write("us,tensor_pct,dram_read_MB,dram_write_MB,grid,kernel\n")
        for d in fwd:
            f.write("%.1f,%.1f,%.1f,%.1f,\"%s\",%s\n" % (
                d.get("gpu__time_duration.sum", 0) / 1e3, d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0),
                d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6, d["grid"], d["name"][:90]))
        tot = sum(d.get("gpu__time_duration.sum", 0) for d in fwd) / 1e3
        tc = sum(d.get("gpu__time_duration.sum", 0) for d in conv) / 1e3
        f.write("# total %.1f us in %d launches; tensor-core conv launches %.1f us (share %.3f)\n" % (tot, len(fwd), tc, tc / tot))
    traffic = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in conv)
    json.dump({"conv_launches": len(conv), "dram_bytes_per_step": traffic, "dram_bytes_per_launch": traffic / max(1, len(conv)),
               "conv_time_us_under_ncu": tc, "share_of_forward_under_ncu": tc / tot, "source": src,
               "how": "dram__bytes_read.sum + dram__bytes_write.sum summed over the tcgen05 conv launches of one forward"},
              open(out_json, "w"), indent=1)
    print(open(out_csv).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
