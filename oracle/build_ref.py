#!/usr/bin/env python3
"""Build the reference's own CPU path into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may load what this script builds.

The reference sources are compiled from where they lie under /root/reference/amg.
Because the shipped tree routes its coarse Krylov SpMV and its interpolation
through CUDA wrappers (with the CPU call commented out directly above each), the
"reference CPU path" only exists after restoring those commented lines (SURVEY.md
Appendix A).  The switches are applied to a throw-away copy in a temp dir outside
the repo; every patched line is asserted against its expected content so a changed
reference fails loudly.  No reference source is ever written into the repo; the
only outputs are shared objects / binaries under oracle/_ref/ (git-ignored).

Outputs:
  oracle/_ref/libsss_ref_fix.so   beta = temp2/temp1 (the author's stated formula; PRIMARY mode)
  oracle/_ref/libsss_ref_asc.so   beta = temp1/temp1 (what the shipped object executes)
  oracle/_ref/libsss_host.a       reference host objects WITHOUT the four solve-phase
                                  objects (SOLVE/cycle/smooth/cuda) -- the link boundary
  oracle/_ref/sss_main.o          the reference's main() (for the drop-in demo binary)
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("AMG_REFERENCE_DIR", "/root/reference/amg")

# (file, 1-based line, must-contain, replacement-or-None(=comment out))
A1 = [
    ("Setup/SSS_inter.cu", 723, "//interp_DIR(A, vertices, P, pars);", "            interp_DIR(A, vertices, P, pars);"),
    ("Setup/SSS_inter.cu", 724, "interp_DIR_cuda(A, vertices, P, pars);", None),
]
# A2: un-comment the CPU SpMV line, comment the *_cuda line below it.
A2_PAIRS = [(91, 92), (169, 172), (259, 260), (314, 315), (390, 391),
            (574, 575), (640, 641), (720, 721), (781, 782)]
# A3: the early "return -1" after failed cudaMalloc (no GPU in the build container).
A3_LINES = [65, 85, 512, 534]
BETA = {
    "fix": ("        beta = temp2 / temp1;", "        temp1 = temp2;"),
    "asc": ("        beta = temp1 / temp1;", "        ;"),
}


def _patch(lines, lineno, must, repl):
    cur = lines[lineno - 1]
    if must not in cur:
        raise SystemExit(f"reference changed: line {lineno}: expected {must!r}, found {cur!r}")
    lines[lineno - 1] = ("// [oracle switch] " + cur.strip()) if repl is None else repl


def stage(mode, tmp):
    dst = os.path.join(tmp, mode)
    os.makedirs(dst)
    for sub in ("", "Setup", "Solve"):
        os.makedirs(os.path.join(dst, sub), exist_ok=True)
        for f in os.listdir(os.path.join(REF, sub)):
            if f.endswith((".c", ".h", ".cu")):
                shutil.copy(os.path.join(REF, sub, f), os.path.join(dst, sub, f))
    # A1
    p = os.path.join(dst, "Setup/SSS_inter.cu")
    L = open(p, encoding="utf-8", errors="surrogateescape").read().split("\n")
    for _, ln, must, repl in A1:
        _patch(L, ln, must, repl)
    open(p, "w", encoding="utf-8", errors="surrogateescape").write("\n".join(L))
    # A2..A4
    p = os.path.join(dst, "Solve/SSS_cycle.cu")
    L = open(p, encoding="utf-8", errors="surrogateescape").read().split("\n")
    for cpu_ln, gpu_ln in A2_PAIRS:
        cur = L[cpu_ln - 1]
        if "//SSS_blas_mv_" not in cur:
            raise SystemExit(f"reference changed: SSS_cycle.cu:{cpu_ln}: {cur!r}")
        L[cpu_ln - 1] = cur.replace("//SSS_blas_mv_", "SSS_blas_mv_", 1)
        _patch(L, gpu_ln, "spmv_cuda(", None)
    for ln in A3_LINES:
        _patch(L, ln, "return -1;", None)
    _patch(L, 373, "beta = temp2_cuda / temp1;", BETA[mode][0])
    _patch(L, 374, "temp1 = temp2_cuda;", BETA[mode][1])
    open(p, "w", encoding="utf-8", errors="surrogateescape").write("\n".join(L))
    return dst


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise SystemExit(f"command failed: {' '.join(cmd)}")


def compile_mode(mode, tmp):
    d = stage(mode, tmp)
    # the reference's own flags (Makefile.sh:3-15) + -fPIC so the objects can live in a .so
    gcc = ["gcc", "-O3", "-fPIC", "-fopenmp", "-c"]
    nvcc = ["nvcc", "-O3", "-Xcompiler", "-fPIC,-fopenmp", "-w", "-c"]
    run(gcc + ["SSS_main.c", "-o", "SSS_main.o"], d)
    run(nvcc + ["SSS_AMG.c", "-o", "SSS_AMG.o"], d)
    run(gcc + ["SSS_matvec.c", "-o", "SSS_matvec.o"], d)
    run(gcc + ["SSS_utils.c", "-o", "SSS_utils.o"], d)
    run(gcc + ["Setup/SSS_coarsen.c", "-o", "SSS_coarsen.o"], d)
    run(nvcc + ["Setup/SSS_SETUP.cu", "-o", "SSS_SETUP.o"], d)
    run(nvcc + ["Setup/SSS_inter.cu", "-o", "SSS_inter.o"], d)
    run(gcc + ["Solve/SSS_SOLVE.c", "-o", "SSS_SOLVE.o"], d)
    run(nvcc + ["Solve/SSS_cycle.cu", "-o", "SSS_cycle.o"], d)
    run(gcc + ["Solve/SSS_smooth.c", "-o", "SSS_smooth.o"], d)
    run(nvcc + ["Solve/SSS_cuda.cu", "-o", "SSS_cuda.o"], d)
    host = ["SSS_AMG.o", "SSS_matvec.o", "SSS_utils.o", "SSS_coarsen.o", "SSS_SETUP.o", "SSS_inter.o"]
    solve = ["SSS_SOLVE.o", "SSS_cycle.o", "SSS_smooth.o", "SSS_cuda.o"]
    so = os.path.join(OUT, f"libsss_ref_{mode}.so")
    run(["nvcc", "-shared", "-Xcompiler", "-fopenmp", "-o", so] + host + solve + ["-lm"], d)
    if mode == "fix":
        lib = os.path.join(OUT, "libsss_host.a")
        if os.path.exists(lib):
            os.remove(lib)
        run(["ar", "rcs", lib] + host, d)
        shutil.copy(os.path.join(d, "SSS_main.o"), os.path.join(OUT, "sss_main.o"))
        bus = os.path.join(OUT, "1138_bus.mtx")
        if os.path.exists(bus):
            os.chmod(bus, 0o644)
        shutil.copy(os.path.join(REF, "Matrix/1138_bus.mtx"), bus)
        os.chmod(bus, 0o644)
        # the drop-in demo: the reference's own main + host objects linked against libamgb200.so
        # (INTEGRATION.md section 1); only possible once the product library has been built
        prod = os.path.join(os.path.dirname(HERE), "amg_b200", "libamgb200.so")
        if os.path.exists(prod):
            exe = os.path.join(OUT, "amg_dropin")
            run(["nvcc", "-Xcompiler", "-fopenmp", "-o", exe, "SSS_main.o"] + host +
                ["-L" + os.path.dirname(prod), "-lamgb200", "-Xlinker", "-rpath=$ORIGIN/../../amg_b200", "-lm"], d)
            print(f"[oracle/_ref] built {os.path.relpath(exe, HERE)} (reference host + libamgb200.so)")
    return so


def main():
    if not os.path.isdir(REF):
        have = os.path.exists(os.path.join(OUT, "libsss_ref_fix.so"))
        print(f"[oracle/_ref] {REF} absent; " + ("using prebuilt oracle/_ref" if have else "NO prebuilt oracle/_ref"))
        return 0 if have else 1
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="amg_ref_build_")
    try:
        for mode in ("fix", "asc"):
            so = compile_mode(mode, tmp)
            print(f"[oracle/_ref] built {os.path.relpath(so, HERE)}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
