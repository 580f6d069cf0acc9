"""Build libpycs_b200.so (sm_100a) in-tree with nvcc.

`python py-cubed-sphere_b200/build.py [--force]` or build() from __graft_entry__.
The operator-surface kernels are compiled with -fmad=false so that each
expression rounds exactly like the reference's numpy expression; the fused
production kernel (fused.cu) uses FMA contraction.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpycs_b200.so")
OBJDIR = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
UNITS = {
    "capi.cu": ["-fmad=false"],
    "halo.cu": ["-fmad=false"],
    "ppm.cu": ["-fmad=false"],
    "wind.cu": ["-fmad=false"],
    "grid.cu": ["-fmad=false"],
    "fused.cu": [],
    "fused2b.cu": [],
    "stepper.cu": [],
    "mgpu.cu": [],
}


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(HERE, "..", "include", "pycs_b200.h"))
    objs, jobs = [], []
    for unit, flags in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJDIR, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((unit, [nvcc] + ARCH + COMMON + flags + ["-c", src, "-o", obj]))

    def compile_one(job):
        unit, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJDIR, unit + ".ptxas.log"), "w") as f:
            f.write(r.stdout + r.stderr)
        return unit, r

    # the translation units are independent: compile them side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1) or 1) as pool:
        for unit, r in pool.map(compile_one, jobs):
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (unit, r.stderr))
            if verbose:
                print(r.stderr)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
