#!/usr/bin/env python
"""Refresh profiles/traffic.json (what bench.py reports as `roofline.traffic`) from an `ncu --set full` summary CSV
made by tools/ncu_all_kernels.sh, stamped with the hash of the CUDA sources the capture was taken on.

    python tools/update_traffic.py gpurun_out/r02_all_kernels.csv [capture-tag]

Only run it when the library the capture profiled was built from the sources in the tree now: bench.py drops the
figures when the hash differs from today's sources.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    # first launch of each: the plain bf16 operator at BASELINE config 2 (tools/ncu_kernels.py order)
    want = {"msda_fwd_pair_kernel": "fwd", "msda_bwd_mma_kernel": "bwd_main", "msda_cvt_f32_bf16_kernel": "bwd_convert"}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        for pat, key in want.items():
            if pat in name and key not in out:
                tot = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(r[ix[m]]) * UNIT[units[ix[m]]]
                out[key] = int(tot)
    out["source_hash"] = bench.library_source_hash()
    out["_source"] = (f"ncu --set full capture {sys.argv[2] if len(sys.argv) > 2 else os.path.basename(path)}, "
                      "dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes")
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(out)


if __name__ == "__main__":
    main()
