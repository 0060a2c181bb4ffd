"""Generate golden vectors from the reference's checked-in generated BVP fixtures.

Run in the authoring container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixture_vectors.py

The reference's fixtures are straight-line three-address Rust (``let tN = ...; out[i] = tN;``) emitted
by its own CodegenIR pipeline for the BVP_Damp1 system (z' = y - z, y' = "-z^3", names [z, y],
z(0)=1, y(N)=1, h = 1, forward scheme) at n_steps = 8 / 32 / 128:
  src/symbolic/codegen/testing_fixtures/test_codegen_generated_bvp_fixtures.rs
  src/symbolic/codegen/testing_fixtures/test_codegen_generated_bvp_large_fixtures.rs
  src/symbolic/codegen/testing_fixtures/test_codegen_generated_bvp_huge_fixtures.rs
This script transliterates each function to Python IN MEMORY (regex; IEEE doubles, ``powf`` -> ``**``),
evaluates it on seeded argument vectors and stores ONLY numbers (args, residual, sparse values) in
tests/golden/bvp_damp1_fixture_vectors.json.  No reference source is copied into the repo.
"""
import json
import math
import os
import re

import numpy as np

REF = "/root/reference/src/symbolic/codegen/testing_fixtures"
FILES = [
    ("test_codegen_generated_bvp_fixtures.rs", "fixture_bvp", 8),
    ("test_codegen_generated_bvp_large_fixtures.rs", "fixture_bvp32", 32),
    ("test_codegen_generated_bvp_huge_fixtures.rs", "fixture_bvp128", 128),
]
FN_RE = re.compile(r"pub fn (\w+)\(args: &\[f64\], out: &mut \[f64\]\) \{(.*?)\n    \}", re.S)


def transliterate(body: str):
    lines = []
    n_out = 0
    for raw in body.splitlines():
        s = raw.strip()
        if not s or s.startswith("debug_assert") or s.startswith("//") or s.startswith('"') or s.startswith(")"):
            continue
        if s.startswith("let "):
            s = s[4:]
        s = s.rstrip(";")
        s = s.replace("_f64", "")
        s = re.sub(r"(\w+)\.powf\((\w+)\)", r"math.pow(\1, \2)", s)
        s = re.sub(r"(\w+)\.exp\(\)", r"math.exp(\1)", s)
        s = re.sub(r"(\w+)\.ln\(\)", r"math.log(\1)", s)
        m = re.match(r"out\[(\d+)\]", s)
        if m:
            n_out = max(n_out, int(m.group(1)) + 1)
        lines.append(s)
    return "\n".join(lines), n_out


def main():
    rng = np.random.default_rng(20260101)
    out = {}
    for fname, prefix, n_steps in FILES:
        src = open(os.path.join(REF, fname)).read()
        fns = {m.group(1): transliterate(m.group(2)) for m in FN_RE.finditer(src)}

        def ordered(kind):
            names = [n for n in fns if n.startswith(f"{prefix}_{kind}_chunk_")]
            return sorted(names, key=lambda n: int(n.rsplit("_", 1)[1]))

        cases = []
        for case in range(3):
            args = (0.2 + 0.9 * rng.random(2 * n_steps)).tolist()
            res, vals = [], []
            for kind, dst in (("residual", res), ("sparse_values", vals)):
                for name in ordered(kind):
                    code, n_out = fns[name]
                    env = {"args": args, "out": [0.0] * n_out, "math": math}
                    exec(code, env)
                    dst.extend(env["out"])
            assert len(res) == 2 * n_steps, (fname, len(res))
            cases.append({"args": args, "residual": res, "sparse_values": vals})
        out[str(n_steps)] = cases
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bvp_damp1_fixture_vectors.json")
    with open(dst, "w") as f:
        json.dump(out, f)
    print("wrote", dst, {k: (len(v[0]["residual"]), len(v[0]["sparse_values"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
