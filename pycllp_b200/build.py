"""Build the CUDA engine in-tree: pycllp_b200/libpycllp_b200.so (sm_100a only).

``python -m pycllp_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpycllp_b200.so")
SOURCES = ["ipm_kernels.cu", "ipm_kernels_py.cu", "ipm_kernels_small.cu", "cabi.cu"]
# per-unit flags: the small-problem kernels are sized for two resident blocks per SM
UNIT_FLAGS = {"ipm_kernels_small.cu": ["-maxrregcount=64"]}
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "-cudart", "shared",
    "--threads", "3",          # the translation units compile side by side
]


OBJDIR = os.path.join(HERE, "build")


def _deps(src):
    """Files a translation unit is built from: itself, this script and the headers it can see
    (cabi.cu only sees *.h; the kernel units see every *.cuh / *.h)."""
    deps = [os.path.join(CSRC, src), os.path.abspath(__file__)]
    exts = (".h",) if src == "cabi.cu" else (".cuh", ".h")
    for d in (CSRC, INCLUDE):
        deps += [os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts)]
    return deps


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile each translation unit to an object (only the stale ones, side by side) and link."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PB200_NVCC_EXTRA", "").split()
    os.makedirs(OBJDIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "--threads", "3")]
    jobs, objs, log = [], [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, _deps(src)):
            cmd = [nvcc] + flags + UNIT_FLAGS.get(src, []) + extra + ["-c", "-o", obj, os.path.join(CSRC, src)]
            jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, proc in jobs:
        out = proc.communicate()[0]
        log.append(" ".join(cmd) + "\n" + out)
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        failed |= proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libpycllp_b200.so")
    if jobs or _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(" ".join(cmd) + "\n" + proc.stdout)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout)
            raise RuntimeError("nvcc failed linking libpycllp_b200.so")
    if log:
        with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as fh:
            fh.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
