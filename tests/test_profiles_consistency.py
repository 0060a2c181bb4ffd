"""The numbers bench.py reports as roofline.traffic come from profiles/ncu_traffic.json, which tools/ncu_traffic.py derives from
the committed ncu launch lists: the committed JSON must be exactly what the tool produces from the committed lists, and the
lists must contain the kernels DESIGN.md names for each workload."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_traffic_json_is_reproducible_from_the_committed_launch_lists(tmp_path):
    work = tmp_path / "repo"
    (work / "tools").mkdir(parents=True)
    (work / "profiles").mkdir()
    shutil.copy(os.path.join(ROOT, "tools", "ncu_traffic.py"), work / "tools")
    for w in ("c2", "c3", "c5"):
        shutil.copy(os.path.join(ROOT, "profiles", f"r02_launches_{w}.csv"), work / "profiles")
    subprocess.run([sys.executable, str(work / "tools" / "ncu_traffic.py")], check=True, capture_output=True)
    got = json.load(open(work / "profiles" / "ncu_traffic.json"))
    want = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert got.keys() == want.keys()
    for key in want:
        for phase in want[key]:
            assert abs(got[key][phase]["bytes_per_interval"] - want[key][phase]["bytes_per_interval"]) < 1e-6, (key, phase)


def test_launch_lists_hold_the_kernels_the_design_names():
    expect = {"c3": ["abd_factor4_kernel<12, 1, 1>", "abd_forward4_kernel<12>", "abd_backward_kernel<12>", "abd_tail_backward_kernel<12>",
                     "assemble_j_staged_kernel", "assemble_f_staged_kernel", "newton_trial_carry_kernel", "newton_reduce_kernel<12>"],
              "c5": ["bigd_factor_kernel", "bigp_forward_kernel", "bigp_backward_kernel", "assemble_wide_kernel"],
              "c2": ["tps_factor_kernel", "tps_forward_kernel", "tps_backward_reduce_kernel", "tps_tail_solve_kernel"]}
    for w, names in expect.items():
        text = open(os.path.join(ROOT, "profiles", f"r02_launches_{w}.csv")).read()
        for n in names:
            assert n in text, (w, n)
