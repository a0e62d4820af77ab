"""Build the native library of astrea_b200 in-tree.

    python -m astrea_b200.build            # nvcc, sm_100a  -> astrea_b200/lib/libastrea_b200.so   (the product)
    python -m astrea_b200.build --hostsim  # g++            -> tests/hostsim/libastrea_hostsim.so  (test infrastructure)

Floating-point flags (DESIGN.md "Numerics"): FMA contraction off and IEEE division / square root, so that the
device arithmetic follows the operation order of the reference numpy code.
"""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "astrea_b200", "csrc")
SOURCES = ["api.cu", "inst_sweep1d.cu", "inst_recon.cu", "inst_flux_pcm.cu", "inst_flux_plm.cu", "inst_flux_ho.cu"]
DEVICE_LIB = os.path.join(ROOT, "astrea_b200", "lib", "libastrea_b200.so")
HOSTSIM_LIB = os.path.join(ROOT, "tests", "hostsim", "libastrea_hostsim.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
GXX_FLAGS = ["-O2", "-std=c++17", "-x", "c++", "-DASTREA_HOSTSIM", "-ffp-contract=off", "-fno-fast-math", "-fPIC",
             "-Wno-unknown-pragmas"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [
        os.path.join(ROOT, "include", "astrea_b200.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), res.stdout, res.stderr))
    return res.stdout + res.stderr


def _compile(cmd, target, verbose):
    out = _run(cmd, verbose)
    os.replace(cmd[-1], target)
    return out


def build(hostsim=False, force=False, verbose=False, ptxas_info=False, jobs=None, variant=None, defines=(), fmad=False):
    """Compile every translation unit (in parallel) and link the shared library.  Returns its path.
    ``variant`` / ``defines``: a build with extra -D flags, written to astrea_b200/lib/variants/<variant>.so (device) or
    tests/hostsim/variants/<variant>.so (host simulation, e.g. the audit build of tests/test_two_pass_audit.py)."""
    lib = HOSTSIM_LIB if hostsim else DEVICE_LIB
    objdir = os.path.join(ROOT, "build", "hostsim" if hostsim else "sm_100a")
    if variant:
        # device tuning builds beside the product library, host-simulated ones beside the test library
        lib = (os.path.join(ROOT, "tests", "hostsim", "variants", variant + ".so") if hostsim
               else os.path.join(ROOT, "astrea_b200", "lib", "variants", variant + ".so"))
        objdir = os.path.join(ROOT, "build", "variant_" + variant)
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    heads = _headers()
    compiler = "g++" if hostsim else os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = list(GXX_FLAGS if hostsim else NVCC_FLAGS) + ["-D" + d for d in defines]
    if fmad and not hostsim:       # tolerance-mode experiments: let ptxas contract a*b+c (results no longer bit-identical)
        flags = [f if f != "-fmad=false" else "-fmad=true" for f in flags]
    if ptxas_info and not hostsim:
        flags += ["-Xptxas", "-v"]
    objs, todo = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + heads + [os.path.abspath(__file__)]):
            todo.append(([compiler] + flags + ["-c", s, "-o", o + ".%d.tmp" % os.getpid()], o))
    logs = []
    if todo:
        with ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as pool:
            logs = list(pool.map(lambda c: _compile(c[0], c[1], verbose), todo))
    if todo or force or _stale(lib, objs):
        tmp = lib + ".%d.tmp" % os.getpid()      # written beside the target and renamed: concurrent builders (pytest-xdist
        if hostsim:                               # workers) never see a half-written library
            _run(["g++", "-shared", "-o", tmp] + objs, verbose)
        else:
            _run([compiler, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], verbose)
        os.replace(tmp, lib)
    if ptxas_info:
        print("\n".join(logs))
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--hostsim", action="store_true")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ptxas-info", action="store_true")
    ap.add_argument("--variant", default=None)
    ap.add_argument("-D", dest="defines", action="append", default=[])
    ap.add_argument("--fmad", action="store_true", help="variant builds only: compile with -fmad=true")
    a = ap.parse_args()
    if a.fmad and not a.variant:
        ap.error("--fmad needs --variant (the product library is built without contraction)")
    print(build(a.hostsim, a.force, a.verbose, a.ptxas_info, variant=a.variant, defines=a.defines, fmad=a.fmad))
    sys.exit(0)
