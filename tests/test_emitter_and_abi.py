"""CPU-side checks of the product's host logic: the CUDA-C emitter, the built library's exported ABI and
the loud failure without a GPU.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import sympy as sp

from oracle import oracle as orc
from rustedscithe_b200 import build as rbuild
from rustedscithe_b200 import lib as rlib
from rustedscithe_b200 import problems
from rustedscithe_b200.emit.cuda_emitter import emit_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return rbuild.build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "rst_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rst_b200_\w+|rustedscithe_aot_\w+)\s*\(", hdr))
    declared -= {"rst_b200_allgather_fn"}
    assert declared == set(rlib.SYMBOLS), declared ^ set(rlib.SYMBOLS)
    L = rlib.load()
    for name in declared:
        assert hasattr(L, name)
    assert L.rst_b200_version() == 1


def test_problem_registry_matches_catalogue(built):
    L = rlib.load()
    keys = [L.rst_b200_problem_key(i).decode() for i in range(L.rst_b200_problem_count())]
    assert keys == rbuild.BUILTIN_PROBLEMS
    for k in keys:
        nv, npar, h = C.c_int(), C.c_int(), C.c_char_p()
        assert L.rst_b200_problem_info(k.encode(), C.byref(nv), C.byref(npar), C.byref(h)) == 0
        spec = problems.get(k)
        assert (nv.value, npar.value) == (spec.nv, len(spec.params))
        assert h.value.decode() == emit_problem(spec).problem_hash


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rustedscithe_b200.solver import NRBVP
    with pytest.raises(rlib.RstB200Error) as e:
        NRBVP("linear2", None, 100)
    assert e.value.code == rlib.ERR_CUDA


def test_create_validates_arguments_before_touching_the_device(built):
    from rustedscithe_b200.solver import NRBVP
    with pytest.raises(rlib.RstB200Error) as e:   # numeric_discretization.rs:60-67
        NRBVP("linear2", None, 100, boundary_conditions={"y": [(0, 0.0)]})
    assert e.value.code == rlib.ERR_INVALID and "boundary conditions" in str(e.value)
    with pytest.raises(rlib.RstB200Error) as e:   # :156-158
        NRBVP("linear2", None, 1)
    assert e.value.code == rlib.ERR_INVALID
    with pytest.raises(KeyError):
        NRBVP("no_such_problem", None, 10)


@pytest.mark.parametrize("key", ["linear2", "damp1", "damp1p", "combustion6", "flame12", "coupled8"])
def test_emitted_expressions_match_sympy_and_the_hand_written_oracle(key):
    """The emitter's pruned df/dy pattern and values agree with (a) direct sympy evaluation of the spec and
    (b) the independently hand-written C right-hand sides of the oracle."""
    spec = problems.get(key)
    em = emit_problem(spec)
    nv = spec.nv
    ys = [sp.Symbol(n, real=True) for n in spec.names]
    ps = [sp.Symbol(n, real=True) for n in spec.params]
    x = sp.Symbol(spec.arg, real=True)
    eqs = [sp.sympify(e) for e in spec.rhs()]
    rng = np.random.default_rng(3)
    yv = 0.3 + 0.5 * rng.random(nv)
    pv = list(spec.param_values)
    subs = dict(zip(ys, yv)) | dict(zip(ps, pv)) | {x: 0.37}
    f_sym = np.array([float(e.evalf(30, subs=subs)) for e in eqs])
    J_sym = np.array([[float(sp.diff(e, y).evalf(30, subs=subs)) for y in ys] for e in eqs])
    f_orc, J_orc = orc.problem_rhs(key, 0.37, yv, pv or None)
    assert np.max(np.abs(f_sym - f_orc) / np.maximum(1e-300, np.abs(f_sym) + 1e-12 * np.max(np.abs(f_sym)))) < 1e-12
    assert np.max(np.abs(J_sym - J_orc) / (np.abs(J_sym) + 1e-12 * np.max(np.abs(J_sym)) + 1e-300)) < 1e-12
    assert np.array_equal(np.array(em.pattern).reshape(nv, nv) != 0, J_orc != 0.0)
    assert np.array_equal(orc.problem_pattern(key), np.array(em.pattern, dtype=np.uint8).reshape(nv, nv))


def test_emitter_rejects_functions_outside_the_reference_opcode_set():
    spec = problems.get("damp1")
    bad = problems.ProblemSpec(key="bad", names=spec.names, rhs=lambda: [sp.gamma(sp.Symbol("z", real=True)), sp.Float(0)],
                               bcs=spec.bcs, bounds=spec.bounds, rtol=spec.rtol, guess=spec.guess)
    with pytest.raises(ValueError, match="opcode"):
        emit_problem(bad)


def test_emitted_literals_round_trip_exactly():
    em = emit_problem(problems.get("combustion6"))
    for lit in re.findall(r"[-+]?\d\.\d{17}e[-+]\d{2}", em.source):
        assert float("%.17e" % float(lit)) == float(lit)


def test_grid_method_numbers_match_the_header():
    """solver.GRID_METHODS (host mirror of GridRefinementMethod, grid_api.rs:62-115) and the RST_B200_GRID_* defines agree."""
    import os
    import re
    from rustedscithe_b200 import solver
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "rst_b200.h")).read()
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define RST_B200_GRID_([A-Z_]+) (\d+)", hdr)}
    names = {"DoublePoints": "DOUBLE_POINTS", "Easy": "EASY", "Pearson": "PEARSON", "GrcarSmooke": "GRCAR_SMOOKE", "Sci": "SCI"}
    assert {names[k]: v for k, v in solver.GRID_METHODS.items()} == defines
    from oracle import adaptive_grid as ag
    assert [ag.DOUBLE_POINTS, ag.EASY, ag.PEARSON, ag.GRCAR_SMOOKE, ag.SCI] == [defines[n] for n in
                                                                               ("DOUBLE_POINTS", "EASY", "PEARSON", "GRCAR_SMOOKE", "SCI")]


@pytest.mark.parametrize("key", ["damp1p", "flame12", "coupled64"])
def test_emitted_cuda_c_text_evaluates_like_the_oracle_on_the_host(key, tmp_path):
    """The emitted struct is plain C++ apart from the CUDA qualifiers: compiled for the host (g++) it must reproduce the
    hand-written oracle right-hand side and Jacobian -- dense `eval_f_fy` for narrow problems, the sink form `eval_f_nz`
    for wide ones -- so the generated TEXT is pinned on the CPU, not only its sympy source."""
    import ctypes
    import subprocess
    spec = problems.get(key)
    em = emit_problem(spec)
    nv = spec.nv
    wide = "eval_f_nz" in em.source
    body = ("HostSink s{J}; P::eval_f_nz(t, y, p, f, s);" if wide else "P::eval_f_fy(t, y, p, f, J);")
    src = tmp_path / f"{key}.cpp"
    src.write_text(
        "#include <cmath>\nusing std::exp; using std::log; using std::sin; using std::cos; using std::tan; using std::pow;\n"
        "using std::sqrt; using std::asin; using std::acos; using std::atan;\n"
        "#define __device__\n#define __forceinline__ inline\n" + em.source +
        f"\nusing P = Problem_{key};\nstruct HostSink {{ double *J; void put(int idx, double v) {{ J[idx] = v; }} }};\n"
        "extern \"C\" void eval_host(double t, const double *y, const double *p, double *f, double *f_only, double *J) {\n"
        f"    for (int i = 0; i < {nv * nv}; ++i) J[i] = 0.0;\n    {body}\n    P::eval_f(t, y, p, f_only);\n}}\n")
    lib = tmp_path / f"{key}.so"
    subprocess.run(["/usr/bin/g++", "-O0", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(lib), str(src)], check=True)
    L = ctypes.CDLL(str(lib))
    dp = ctypes.POINTER(ctypes.c_double)
    L.eval_host.argtypes = [ctypes.c_double, dp, dp, dp, dp, dp]
    rng = np.random.default_rng(8)
    yv = np.ascontiguousarray(0.3 + 0.5 * rng.random(nv))
    pv = np.ascontiguousarray(list(spec.param_values) or [0.0], dtype=np.float64)
    f, f_only, J = np.zeros(nv), np.zeros(nv), np.zeros(nv * nv)
    L.eval_host(0.37, yv.ctypes.data_as(dp), pv.ctypes.data_as(dp), f.ctypes.data_as(dp), f_only.ctypes.data_as(dp),
                J.ctypes.data_as(dp))
    f_orc, J_orc = orc.problem_rhs(key, 0.37, yv, list(spec.param_values) or None)
    scale_f = np.abs(f_orc) + 1e-12 * np.max(np.abs(f_orc)) + 1e-300
    assert np.max(np.abs(f - f_orc) / scale_f) < 1e-12 and np.max(np.abs(f_only - f_orc) / scale_f) < 1e-12
    J = J.reshape(nv, nv)
    assert np.max(np.abs(J - J_orc) / (np.abs(J_orc) + 1e-12 * np.max(np.abs(J_orc)) + 1e-300)) < 1e-12
    assert np.array_equal(J != 0.0, np.array(em.pattern).reshape(nv, nv) != 0)


def test_manifest_h_matches_the_reference_format_and_the_oracle_structure():
    """f3: manifest.h as the reference's C AOT library writer emits it (codegen_c_aot_library.rs:250-278), with the entry
    count of the oracle's BandedJacobianStructure."""
    from oracle import oracle as orc
    from rustedscithe_b200 import build as rbuild
    from rustedscithe_b200 import problems
    for key, n, scheme in (("linear2", 80, "forward"), ("flame12", 33, "forward"), ("combustion6", 50, "trapezoid"),
                           ("coupled64", 7, "forward")):
        spec = problems.get(key)
        o = orc.OracleBVP(key, n, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, scheme=scheme)
        rows = o.structure()[0]
        text = rbuild.manifest_h(spec, n, scheme)
        nn = spec.nv * n
        assert text == ("/* AUTO-GENERATED C AOT MANIFEST */\n\n#ifndef MANIFEST_H\n#define MANIFEST_H\n\n"
                        "#define BACKEND_KIND \"aot\"\n#define MATRIX_BACKEND \"banded\"\n"
                        f"#define RESIDUAL_LEN {nn}\n#define JACOBIAN_ROWS {nn}\n#define JACOBIAN_COLS {nn}\n"
                        f"#define JACOBIAN_NNZ {rows.shape[0]}\n"
                        "#define RESIDUAL_FN_NAME \"eval_residual\"\n#define JACOBIAN_FN_NAME \"eval_banded_values\"\n\n"
                        "#endif /* MANIFEST_H */\n")
