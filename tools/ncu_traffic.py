"""Build profiles/ncu_traffic.json from the committed ncu launch lists (profiles/r02_launches_*.csv).

Each list is `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`
over `bench.py --workload <w> --steps K --warmup 3 --no-cpu --no-secondary`.  For every phase bench.py times
(assemble_FJ, assemble_F, factor, solve) the DRAM bytes of ALL its kernels are summed and divided by the number of times the
phase ran (= launches of its level-0 kernel) and by the mesh size: bytes per interval, the figure `roofline.traffic` scales.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LISTS = {"flame12": ("profiles/r02_launches_c3.csv", 1_000_000), "linear2": ("profiles/r02_launches_c2.csv", 1_000_000),
         "coupled64": ("profiles/r02_launches_c5.csv", 250_000)}


def phase_of(name):
    if re.search(r"assemble_j_staged|assemble_kernel<[^,]+, *\(bool\)1|assemble_kernel<[^,]+, 1|assemble_wide_kernel<[^,]+, *(\(bool\))?1", name):
        return "assemble_FJ"
    if re.search(r"assemble_f_staged|assemble_kernel|assemble_wide_kernel", name):
        return "assemble_F"
    if "factor" in name:
        return "factor"
    if re.search(r"forward|backward|top_solve|tail_solve", name):
        return "solve"
    return None


def main():
    out = {}
    for key, (path, n) in LISTS.items():
        rows = list(csv.reader(open(os.path.join(ROOT, path))))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        idx = {h: i for i, h in enumerate(rows[hi])}
        per = collections.defaultdict(lambda: collections.defaultdict(float))    # phase -> (kernel, grid) -> bytes / launches
        for r in rows[hi + 1:]:
            if len(r) < len(rows[hi]):
                continue
            name, grid, metric = r[idx["Kernel Name"]], r[idx["Grid Size"]], r[idx["Metric Name"]]
            ph = phase_of(name)
            if ph is None:
                continue
            v = float(r[idx["Metric Value"]].replace(",", ""))
            k = (name.split("(")[0], grid)
            if metric.startswith("dram__bytes"):
                per[ph][k] += v
                per[ph][("bytes", "total")] += v
            elif metric.startswith("gpu__time"):
                per[ph][(k, "launches")] += 1
        out[key] = {}
        for ph, d in per.items():
            launches = {k[0]: v for k, v in d.items() if isinstance(k[0], tuple)}
            # the phase ran as often as its biggest-grid kernel was launched
            big = max(launches, key=lambda kg: int(kg[1].strip("()").split(",")[0]))
            runs = launches[big]
            out[key][ph] = {"bytes_per_interval": d[("bytes", "total")] / runs / n, "phase_runs_in_capture": int(runs),
                            "source": f"{path}: dram__bytes_read.sum + dram__bytes_write.sum of every kernel of the phase, "
                                      f"per phase execution, / {n} intervals"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    for key, d in out.items():
        print(key, {ph: round(v["bytes_per_interval"], 1) for ph, v in d.items()})


if __name__ == "__main__":
    main()
