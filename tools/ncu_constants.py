#!/usr/bin/env python
"""Refresh the ncu-derived constants of one scene in tools/algo_bytes.json (instructions per sample, lanes per
instruction, DRAM bytes per sample) from an .ncu-rep of ONE launch of that scene.
    python tools/ncu_constants.py <report.ncu-rep> <scene> <samples in the launch> [source note]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, scene, samples = sys.argv[1], sys.argv[2], float(sys.argv[3])
    note = sys.argv[4] if len(sys.argv) > 4 else None
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-kernel-base", "mangled"], text=True,
                                  stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def num(key):
        i = hdr.index(key)
        v = float(vals[i].replace(",", ""))
        unit = units[i].lower()
        return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)

    path = os.path.join(ROOT, "tools", "algo_bytes.json")
    doc = json.load(open(path))
    ncu = doc["scenes"][scene].setdefault("ncu", {})
    ncu["kernel"] = subprocess.check_output(["c++filt", vals[hdr.index("Kernel Name")]], text=True).strip()   # with its namespace
    ncu["inst_per_sample"] = num("smsp__inst_executed.sum") / samples
    ncu["lanes"] = num("smsp__thread_inst_executed_per_inst_executed.ratio")
    ncu["dram_bytes_per_sample"] = (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / samples
    if note:
        ncu["source"] = note
    json.dump(doc, open(path, "w"), indent=1)
    print(scene, json.dumps(ncu))


if __name__ == "__main__":
    main()
