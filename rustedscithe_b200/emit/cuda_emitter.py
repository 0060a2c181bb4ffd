"""CUDA-C emitter for the AOT residual/Jacobian pipeline (north_star subsystem 1).

Role in the reference: ``src/symbolic/codegen/c_backend/c_emitter.rs:265-344`` lowers the symbolic
residual rows and the zero-pruned sparse Jacobian entries (``src/symbolic/View/jacobian.rs:95-137``)
through ``CodegenIR`` (17 opcodes, CSE cache ``CodegenIR.rs:1043-1060``) into three-address C, one
straight-line expression per *discretised* row -- fully unrolled over the mesh, which is what stops it
at ~10^4 points.

This emitter is the 4th sibling of that C emitter but lowers the *continuous* right-hand side
``f(t, y, p)`` and its symbolically differentiated, zero-pruned Jacobian ``df/dy`` ONCE, as a
``__device__`` function evaluated per interval with a run-time interval index.  The mesh loop, the
discretisation arithmetic (``symbolic_functions_BVP.rs:5486-5490``) and the block layout live in
``csrc/assemble.cuh``; the generated struct only supplies the per-point mathematics.

Pipeline: sympy expressions -> ``sympy.diff`` per (row, variable) -> exact-zero pruning -> ``sympy.cse``
over f and df/dy together -> straight-line IR restricted to the reference's opcode set (const, input,
+ - * /, pow, exp, ln, sin, cos, tan, cot, asin, acos, atan, acot) -> CUDA C.  Literals are printed
with 17 significant digits (``c_emitter.rs:59-70`` uses ``{:.17e}``) so constants round-trip exactly.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import sympy as sp
from sympy.printing.c import C99CodePrinter

WIDE_NV = 16   # above this block size the Jacobian is emitted through a sink (eval_f_nz) instead of a dense fy[]
ALLOWED_FUNCS = {"exp", "log", "sin", "cos", "tan", "cot", "asin", "acos", "atan", "acot"}


class _CudaPrinter(C99CodePrinter):
    """C99 printer with exact 17-digit literals and multiplication chains for small integer powers."""

    def _print_Float(self, expr):
        return "%.17e" % float(expr)

    def _print_Integer(self, expr):
        return "%.17e" % float(int(expr))

    def _print_Rational(self, expr):
        return "(%.17e/%.17e)" % (float(expr.p), float(expr.q))

    def _print_Pow(self, expr):
        base, exp = expr.as_base_exp()
        if exp.is_Integer and 2 <= int(exp) <= 4:
            b = self._print(base)
            if not base.is_Atom:
                b = f"({b})"
            return "(" + "*".join([b] * int(exp)) + ")"
        if exp.is_Integer and -4 <= int(exp) <= -1:
            b = self._print(base)
            if not base.is_Atom:
                b = f"({b})"
            return "(1.0/(" + "*".join([b] * (-int(exp))) + "))"
        if exp == sp.Rational(1, 2):
            return f"sqrt({self._print(base)})"
        return f"pow({self._print(base)}, {self._print(exp)})"

    def _print_cot(self, expr):
        return f"(1.0/tan({self._print(expr.args[0])}))"

    def _print_acot(self, expr):
        return f"atan(1.0/({self._print(expr.args[0])}))"


@dataclass
class EmittedProblem:
    key: str
    nv: int
    n_params: int
    pattern: list          # nv*nv 0/1, structural non-zeros of df/dy (after zero pruning)
    source: str            # CUDA-C text of the struct
    problem_hash: str      # analogue of the reference's problem_key (codegen_aot_registry.rs)
    n_temps: int


def _check_ops(expr):
    for node in sp.preorder_traversal(expr):
        if isinstance(node, sp.Function):
            name = node.func.__name__
            if name not in ALLOWED_FUNCS:
                raise ValueError(f"function '{name}' is outside the reference's codegen opcode set")


def emit_problem(spec) -> EmittedProblem:
    """Lower ``spec`` (rustedscithe_b200.problems.ProblemSpec) to a CUDA-C device struct."""
    names = list(spec.names)
    nv = len(names)
    ysyms = [sp.Symbol(n, real=True) for n in names]
    psyms = [sp.Symbol(n, real=True) for n in spec.params]
    tsym = sp.Symbol(spec.arg, real=True)
    eqs = [sp.sympify(e) for e in spec.rhs()]
    if len(eqs) != nv:
        raise ValueError("number of equations must equal the number of variables")
    for e in eqs:
        _check_ops(e)
        extra = e.free_symbols - set(ysyms) - set(psyms) - {tsym}
        if extra:
            raise ValueError(f"unknown symbols in right-hand side: {extra}")

    # symbolic differentiation restricted to variables that occur in the row (zero pruning,
    # View/jacobian.rs:111-128), exact zeros dropped
    jac = {}
    pattern = [0] * (nv * nv)
    for i, e in enumerate(eqs):
        fs = e.free_symbols
        for c, ys in enumerate(ysyms):
            if ys not in fs:
                continue
            d = sp.diff(e, ys)
            if d == 0:
                continue
            jac[(i, c)] = d
            pattern[i * nv + c] = 1

    printer = _CudaPrinter()
    subs = {ys: sp.Symbol(f"y[{k}]") for k, ys in enumerate(ysyms)}
    subs.update({ps: sp.Symbol(f"p[{k}]") for k, ps in enumerate(psyms)})
    subs[tsym] = sp.Symbol("t")

    def lower(exprs, out_names):
        reps, red = sp.cse([ex.xreplace(subs) for ex in exprs], symbols=sp.numbered_symbols("w"), order="none")
        lines = [f"        const double {printer.doprint(s)} = {printer.doprint(v)};" for s, v in reps]
        lines += [f"        {o} = {printer.doprint(v)};" for o, v in zip(out_names, red)]
        return lines, len(reps)

    f_lines, nt_f = lower(eqs, [f"f[{i}]" for i in range(nv)])
    jkeys = sorted(jac)
    fj_lines, nt_fj = lower(eqs + [jac[k] for k in jkeys],
                            [f"f[{i}]" for i in range(nv)] + [f"fy[{i * nv + c}]" for (i, c) in jkeys])
    zero_lines = [f"        fy[{i * nv + c}] = 0.0;" for i in range(nv) for c in range(nv) if not pattern[i * nv + c]]

    cls = f"Problem_{spec.key}"
    pat_txt = ", ".join(str(v) for v in pattern)
    src = []
    src.append(f"// generated by rustedscithe_b200/emit/cuda_emitter.py -- do not edit")
    src.append(f"// problem '{spec.key}': variables {names}, parameters {list(spec.params)}, argument '{spec.arg}'")
    src.append(f"struct {cls} {{")
    src.append(f"    static constexpr int NV = {nv};")
    src.append(f"    static constexpr int NP = {len(psyms)};")
    src.append(f"    static constexpr int NNZ = {len(jkeys)};")
    src.append(f"    static const char *key() {{ return \"{spec.key}\"; }}")
    src.append(f"    static const unsigned char *pattern() {{ static const unsigned char pat[{nv * nv}] = {{{pat_txt}}}; return pat; }}")
    src.append("    // f(t, y, p)")
    src.append("    __device__ __forceinline__ static void eval_f(double t, const double *y, const double *p, double *f) {")
    src.append("        (void)t; (void)y; (void)p;")
    src += f_lines
    src.append("    }")
    if nv <= WIDE_NV:
        src.append("    // f and the zero-pruned df/dy (row-major NV*NV, structural zeros written as 0.0)")
        src.append("    __device__ __forceinline__ static void eval_f_fy(double t, const double *y, const double *p, double *f, double *fy) {")
        src.append("        (void)t; (void)y; (void)p;")
        src += zero_lines
        src += fj_lines
        src.append("    }")
    else:
        # wide blocks: a dense per-thread fy[NV*NV] would live in local memory; the structural non-zeros are handed to a
        # sink (the assembly kernel stores them straight into the pre-initialised block, csrc/assemble.cuh)
        src.append("    // f and the structural non-zeros of df/dy: sink.put(row * NV + col, value), each entry exactly once")
        src.append("    template <class Sink>")
        src.append("    __device__ __forceinline__ static void eval_f_nz(double t, const double *y, const double *p, double *f, Sink &sink) {")
        src.append("        (void)t; (void)y; (void)p;")
        src += [ln.replace("        fy[", "        sink.put(").replace("] = ", ", ", 1).replace(";", ");")
                if ln.startswith("        fy[") else ln for ln in fj_lines]
        src.append("    }")
    src.append("};")
    text = "\n".join(src) + "\n"
    return EmittedProblem(spec.key, nv, len(psyms), pattern, text,
                          hashlib.sha256(text.encode()).hexdigest()[:16], max(nt_f, nt_fj))


def emit_registry(emitted, path_header: str, path_registry: str):
    """Write ``generated_problems.cuh`` (the structs) and ``generated_registry.inc`` (the X-macro list
    the core library instantiates kernels from)."""
    with open(path_header, "w") as f:
        f.write("// generated by rustedscithe_b200/emit -- do not edit\n#pragma once\n\n")
        for e in emitted:
            f.write(e.source)
            f.write("\n")
    with open(path_registry, "w") as f:
        f.write("// generated by rustedscithe_b200/emit -- do not edit\n")
        for e in emitted:
            f.write(f"RST_PROBLEM(Problem_{e.key}, \"{e.problem_hash}\")\n")
