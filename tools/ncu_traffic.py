"""Map the kernels of an ncu --set full capture of `tools/bench_kernels.py --modes ... --iters 1 --warm 0` onto bench.py's
kernel tags (in launch order) and write profiles/ncu_traffic.json = {tag: dram bytes per launch}."""
import csv
import io
import json
import subprocess
import sys

rep, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
order = sys.argv[4].split(",")  # e.g. "f32:zdown2,f32:up2,f64:defect"
TAGS = {"smooth2": "rbgs2", "zdown2": "Z+rbgs2+R", "down2": "rbgs2+R", "up2": "P+rbgs2", "up2norm": "P+rbgs2+N",
        "defect": "update+resid32+N", "jac2": "jac2", "jacdown2": "jac2+R", "jacup2": "P+jac2"}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


res = {}
for spec, r in zip(order, data):
    dt, mode = spec.split(":")
    rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    res[f"{TAGS[mode]}/{dt}/{n}x{n}"] = rd + wr
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
