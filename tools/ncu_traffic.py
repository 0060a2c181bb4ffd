"""Build profiles/ncu_traffic.json from the committed ncu launch lists (profiles/r02_launches_*.csv).

Each list is `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`
over `bench.py --workload <w> --steps K --warmup 3 --no-cpu --no-secondary`.  For every phase bench.py times
(assemble_FJ, assemble_F, factor, solve) the DRAM bytes of ALL its kernels are summed and divided by the number of times the
phase ran and by the mesh size: bytes per interval, the figure `roofline.traffic` scales.
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
        # phase -> (kernel, grid) -> [bytes, launches].  A phase execution launches every (kernel, grid) of the phase once (one
        # launch per level and sweep direction), so its traffic is the sum over them of the mean bytes per launch -- robust
        # against phases that skip kernels (the first solve after a refactorisation runs no forward sweeps).  The two variants of
        # the nv = 12 factor kernel (with / without the right-hand-side column) count as one kernel.
        per = collections.defaultdict(lambda: collections.defaultdict(lambda: [0.0, 0]))
        for r in rows[hi + 1:]:
            if len(r) < len(rows[hi]):
                continue
            name, grid, metric = r[idx["Kernel Name"]], r[idx["Grid Size"]], r[idx["Metric Name"]]
            ph = phase_of(name)
            if ph is None:
                continue
            v = float(r[idx["Metric Value"]].replace(",", ""))
            kname = re.sub(r"(abd_factor4_kernel<\d+, \d), \d>", r"\1>", name.split("(")[0])
            kname = kname.replace("tps_backward_reduce_kernel", "tps_backward_kernel")   # level-0 sweep with / without the fused reductions
            k = (kname, grid)
            if metric.startswith("dram__bytes"):
                per[ph][k][0] += v
            elif metric.startswith("gpu__time"):
                per[ph][k][1] += 1
        out[key] = {}
        for ph, d in per.items():
            total = sum(b / max(c, 1) for b, c in d.values())
            out[key][ph] = {"bytes_per_interval": total / n, "kernels_per_phase_execution": len(d),
                            "source": f"{path}: dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of every (kernel, grid) "
                                      f"of the phase, summed, / {n} intervals"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    for key, d in out.items():
        print(key, {ph: round(v["bytes_per_interval"], 1) for ph, v in d.items()})


if __name__ == "__main__":
    main()
