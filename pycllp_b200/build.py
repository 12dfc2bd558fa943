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
SOURCES = ["ipm_kernels.cu", "ipm_kernels_py.cu", "cabi.cu"]
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "-cudart", "shared",
    "--threads", "3",          # the translation units compile side by side
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    # every source and header the library is built from (csrc/*.cu, *.cuh, *.h, include/*.h)
    deps = [os.path.abspath(__file__)]
    for d in (CSRC, INCLUDE):
        deps += [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cu", ".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PB200_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libpycllp_b200.so")
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + proc.stdout)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
