"""Build the B200 library: run the CUDA-C emitter over the problem catalogue, then nvcc for sm_100a.

``python -m rustedscithe_b200.build`` (also called by ``__graft_entry__.build()``).  The library is built
IN-TREE (rustedscithe_b200/librst_b200.so) so it travels to the GPU box with the snapshot.  This is the
analogue of the reference's AOT build step (src/symbolic/codegen/c_backend/codegen_c_aot_build.rs:44-84
probes and invokes gcc/tcc; here the compiler is nvcc and the target is sm_100a only).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "generated")
LIB = os.path.join(HERE, "librst_b200.so")
BUILTIN_PROBLEMS = ["linear2", "damp1", "damp1p", "combustion6", "flame12", "coupled8"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def emit_sources() -> None:
    from . import problems
    from .emit.cuda_emitter import emit_problem, emit_registry

    os.makedirs(GEN, exist_ok=True)
    emitted = [emit_problem(problems.get(k)) for k in BUILTIN_PROBLEMS]
    hdr, reg = os.path.join(GEN, "generated_problems.cuh"), os.path.join(GEN, "generated_registry.inc")
    tmp_h, tmp_r = hdr + ".tmp", reg + ".tmp"
    emit_registry(emitted, tmp_h, tmp_r)
    for tmp, dst in ((tmp_h, hdr), (tmp_r, reg)):  # keep mtimes stable when nothing changed
        if os.path.exists(dst) and open(dst).read() == open(tmp).read():
            os.remove(tmp)
        else:
            os.replace(tmp, dst)


def _sources():
    out = [os.path.join(HERE, "..", "include", "rst_b200.h")]
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".inc", ".h"))]
    return sorted(out)


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in _sources():
        h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    emit_sources()
    stamp = LIB + ".stamp"
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", LIB, os.path.join(CSRC, "rst_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building librst_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
