"""DRAM traffic of the gwd_tapgemm_kernel launches of one step from an ncu CSV
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gwd_tapgemm ...
-> JSON (bench.py reads profiles/r2_tapgemm_dram_train.json for `roofline.traffic`).
usage: python tools/summarize_dram.py <csv> <out.json> "<how>" """
import csv
import json
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)


def main(path, out, how):
    rows = [ln for ln in open(path) if ln.startswith('"')]
    rd = wr = us = 0.0
    ids = set()
    for r in csv.DictReader(rows):
        if "gwd_tapgemm_kernel" not in r["Kernel Name"]:
            continue
        ids.add(r["ID"])
        if r["Metric Name"] == "dram__bytes_read.sum":
            rd += to_bytes(r["Metric Value"], r["Metric Unit"])
        elif r["Metric Name"] == "dram__bytes_write.sum":
            wr += to_bytes(r["Metric Value"], r["Metric Unit"])
        elif r["Metric Name"] == "gpu__time_duration.sum":
            us += to_us(r["Metric Value"], r["Metric Unit"])
    n = len(ids)
    d = {"kernel": "gwd_tapgemm_kernel", "launches": n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
         "dram_bytes_per_launch": (rd + wr) / max(n, 1), "kernel_us_per_step_under_ncu": us, "how": how}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
