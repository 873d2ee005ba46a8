#!/usr/bin/env python
"""Regenerates profiles/<round>_traffic.json and profiles/<round>_launches.csv on the GPU box:

    python profiles/capture_traffic.py r02        (under gpurun; ~2 min)

Runs bench.py (the default config-3 workload, few steps) under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`
and averages DRAM bytes (read + write) and duration per launch for every kernel of the library.
bench.py reads the JSON for its `roofline.traffic` / `roofline_k_filter.moved_bytes` fields; numbers
printed by the profiled run itself are never used as bench values."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    raw = os.path.join(out_dir, f"{tag}_launches_raw.csv")
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-c", "600", "--csv", "--log-file", raw, sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
           "--no-cpu", "--no-ref-gpu", "--no-streamed", "--no-latency", "--prod-iters", "2", "--dense-iters", "1"] + sys.argv[2:]
    subprocess.run(cmd, check=True, stdout=open(os.path.join(out_dir, f"{tag}_traffic_bench.log"), "w"), stderr=subprocess.STDOUT)
    rows = [r for r in csv.reader(open(raw)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = {n: i for i, n in enumerate(rows[hdr])}
    per = defaultdict(lambda: defaultdict(list))  # kernel -> launch id -> metrics
    order = []
    for r in rows[hdr + 1:]:
        if len(r) <= cols["Metric Value"]:
            continue
        name = r[cols["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
        name = name[:-2] if name.endswith("_t") else name  # k_filter_t<...> -> k_filter
        lid = r[cols["ID"]]
        unit, val = r[cols["Metric Unit"]], float(r[cols["Metric Value"]].replace(",", ""))
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(unit, 1.0)
        per[name][lid].append((r[cols["Metric Name"]], val * scale))
        if (name, lid) not in order:
            order.append((name, lid))
    bytes_per, ns_per, count = {}, {}, {}
    for name, launches in per.items():
        b, t = [], []
        for lid, ms in launches.items():
            d = dict(ms)
            b.append(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0))
            t.append(d.get("gpu__time_duration.sum", 0.0))
        # the production kernels: skip the first (cold, table rebuild) launches when there are enough
        skip = 2 if len(b) > 4 else 0
        bytes_per[name] = sum(b[skip:]) / max(1, len(b[skip:]))
        ns_per[name] = sum(t[skip:]) / max(1, len(t[skip:]))
        count[name] = len(b)
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                         "on `bench.py --steps 2 --warmup 3` (profiles/capture_traffic.py); mean per launch, first two launches "
                         "of a kernel skipped; per-launch times under ncu are cold-cache and serialised",
               "bytes_per_launch": bytes_per, "us_per_launch_under_ncu": {k: v / 1e3 for k, v in ns_per.items()}, "launches": count},
              open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        f.write("kernel,launch_id,gpu_time_ns,dram_read_bytes,dram_write_bytes\n")
        for name, lid in order:
            d = dict(per[name][lid])
            f.write(f"{name},{lid},{d.get('gpu__time_duration.sum', 0):.0f},{d.get('dram__bytes_read.sum', 0):.0f},"
                    f"{d.get('dram__bytes_write.sum', 0):.0f}\n")
    # the files must travel back: gpurun only merges gpurun_out/
    for n in (f"{tag}_traffic.json", f"{tag}_launches.csv"):
        subprocess.run(["cp", os.path.join(ROOT, "profiles", n), os.path.join(out_dir, n)], check=True)
    print(json.dumps({"kernels": count, "bytes_per_launch": bytes_per}))


if __name__ == "__main__":
    main()
