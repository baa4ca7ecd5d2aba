"""Collect ncu-measured DRAM traffic per launch into profiles/ncu_traffic.json (read by bench.py for roofline.traffic).

    python profiles/make_traffic.py profiles/r01c_step_full_raw.csv:step profiles/r01h_launches_bw.csv:bandwidth

Inputs are ncu CSV exports committed under profiles/:  `--page raw --csv` of a `--set full` capture (one row per
launch, metrics as columns) or the `--csv --log-file` launch list of a `--metrics ...` pass (one row per metric).
traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the last captured launch of each kernel:
warm caches, like the timed region)."""
import csv
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.split("(")[0].split("<")[0]
    return name.split("::")[-1]


def wide(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        def get(m, table):
            if m not in col or r[col[m]] in ("", "n/a"):
                return None
            return float(r[col[m]].replace(",", "")) * table[units[col[m]]]
        yield short(r[col["Kernel Name"]]), get("dram__bytes_read.sum", UNIT), get("dram__bytes_write.sum", UNIT), \
            get("gpu__time_duration.sum", TIME)


def tall(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {n: i for i, n in enumerate(rows[h])}
    cur = {}
    for r in rows[h + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        d = cur.setdefault(r[col["ID"]], {"name": short(r[col["Kernel Name"]])})
        table = UNIT if "bytes" in r[col["Metric Name"]] else TIME
        d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", "")) * table.get(r[col["Metric Unit"]], 1.0)
    for d in cur.values():
        yield d["name"], d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), d.get("gpu__time_duration.sum")


def main():
    out = {}
    for arg in sys.argv[1:]:
        path, _, regime = arg.partition(":")
        first = open(path).readline()
        it = wide(path) if first.startswith('"ID"') else tall(path)
        for name, rd, wr, us in it:
            if rd is None or wr is None or not name:
                continue
            prev = out.setdefault(regime or "step", {}).get(name)
            if prev is not None and prev["dram_bytes_per_launch"] >= rd + wr:
                continue                      # several launches of one kernel: keep the largest (the uniform-id case)
            out[regime or "step"][name] = {
                "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                "duration_us_under_ncu": None if us is None else round(us, 2), "source": os.path.relpath(path)}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
