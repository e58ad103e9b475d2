"""DRAM bytes per launch of every nif_* kernel from an `ncu --page raw --csv` export (what bench.py's roofline.traffic reads).

    python tools/ncu_kernels_json.py gpurun_out/X_raw.csv profiles/r02x_ncu_kernels.json [batch]
"""
import csv
import json
import re
import sys

r = list(csv.reader(open(sys.argv[1])))
hdr, units, rows = r[0], r[1], r[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for row in rows:
    name = re.sub(r"^void\s+", "", row[ix["Kernel Name"]])
    name = re.split(r"[<(]", name)[0]
    if not name.startswith("nif_"):
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(row[ix[m]].replace(",", "")) * scale[units[ix[m]]]
    acc.setdefault(name, []).append(tot)
out = {k: sum(v) / len(v) for k, v in acc.items()}  # mean over the launches of the capture
if len(sys.argv) > 3:
    out["_batch"] = int(sys.argv[3])  # rows per step of the captured run (bench.py only uses a capture of its own batch)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
