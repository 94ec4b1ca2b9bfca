#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full --page raw --csv` export of ONE step of the default bench workload:
DRAM bytes (read + write) per launch group, keyed by a hash of the kernel sources it was captured from.
Usage: python scripts/make_traffic_json.py profiles/r2_ncu_full_raw.csv "<what the capture was>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_sha  # noqa: E402

GROUP = {"k_unstuff": "unstuff", "k_subseq_table": "unstuff", "k_huff_sync": "sync", "k_huff_write": "write", "k_zero_tail": "write",
         "k_dc_predict": "write", "k_idct_color": "idct"}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {"unstuff": 0.0, "sync": 0.0, "write": 0.0, "idct": 0.0}
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").split("<")[0].strip()
        g = GROUP.get(name)
        if g:
            out[g] += float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
    doc = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch group of one step, bytes; bench.py reports it as roofline.traffic "
                       "only while kernel_sources_sha and the workload match the run",
           "capture": sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]),
           "kernel_sources_sha": kernel_sources_sha(), "workload": "config2", "images": 4096}
    doc.update({k: int(v) for k, v in out.items()})
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc))


if __name__ == "__main__":
    main()
