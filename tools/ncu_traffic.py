#!/usr/bin/env python
"""Developer tool: DRAM traffic per launch of one kernel from an `ncu --page raw --csv` export -> profiles/traffic.json.

    ncu -i gpurun_out/rNN/ncu_full_<k>.ncu-rep --page raw --csv > profiles/rNN/ncu_full_raw_<k>.csv
    python tools/ncu_traffic.py profiles/rNN/ncu_full_raw_<k>.csv c2 'tile_kernel.*JsdOp'

The entry records the hash of csrc/ it was captured from (`csrc_sha16`, the same hash bench.py computes): bench.py copies an
entry into `roofline.traffic` only while the kernels are still the ones that were profiled.
"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "deep-co-training-for-semi-supervised-image-segmentation_b200", "csrc")


def csrc_sha16():
    """Hash of the kernel sources the loaded library was built from (tools/ncu_traffic.py stamps captures with it).
    Comments and white space are stripped first: only a change that can alter the generated code invalidates a capture."""
    import hashlib
    import re
    d = CSRC
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            src = open(os.path.join(d, f), "r", encoding="utf-8", errors="replace").read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            src = re.sub(r"//[^\n]*", "", src)
            h.update(f.encode())
            h.update("".join(src.split()).encode())
    return h.hexdigest()[:16]


def main():
    path, workload, pattern = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
    rows = list(csv.reader(open(path, newline="")))
    head = rows[0]
    col = {n: i for i, n in enumerate(head)}
    rd = next(i for n, i in col.items() if n.endswith("dram__bytes_read.sum"))
    wr = next(i for n, i in col.items() if n.endswith("dram__bytes_write.sum"))
    units = rows[1] if rows[1][col["ID"]] == "" else None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    su_r = scale.get(units[rd], 1.0) if units else 1.0
    su_w = scale.get(units[wr], 1.0) if units else 1.0
    hits = [r for r in rows[1:] if len(r) > wr and pattern.search(r[col["Kernel Name"]])]
    if not hits:
        sys.exit(f"no launch of /{pattern.pattern}/ in {path}")
    f = lambda s: float(s.replace(",", ""))  # noqa: E731
    read = sum(f(r[rd]) for r in hits) * su_r / len(hits)
    write = sum(f(r[wr]) for r in hits) * su_w / len(hits)
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(tj)) if os.path.exists(tj) else {}
    data[workload] = {"bytes": int(read + write), "read": int(read), "write": int(write), "launches": len(hits),
                      "kernel": hits[0][col["Kernel Name"]][:120], "csrc_sha16": csrc_sha16(),
                      "source": os.path.relpath(path, ROOT)}
    json.dump(data, open(tj, "w"), indent=2)
    print(json.dumps(data[workload], indent=1))


if __name__ == "__main__":
    main()
