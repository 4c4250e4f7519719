"""Build recipe of libo3d_b200.so: explicit nvcc for sm_100a, in-tree output.

    python -m osinco3d_b200.build [--force] [--verbose]

The library is plain CUDA C++ with a C ABI (include/o3d_b200.h): no torch, no pybind.
-fmad=false: FMA contraction is disabled on purpose so the FP64 stencil arithmetic is
bit-identical to the reference's gfortran x86-64 build (see DESIGN.md "Arithmetic").
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(OUT_DIR, "libo3d_b200.so")

SOURCES = ["api.cu", "modules.cu", "pipeline.cu", "poisson.cu", "comm.cu", "tma.cu", "ghost_kernels.cu",
           "der_kernel.cu", "vel_kernels.cu", "proj_kernels.cu", "sor_kernels.cu", "sor_tma_kernel.cu", "sor_persist_kernel.cu",
           "transeq_kernels.cu", "reduce_kernels.cu", "mg_kernels.cu", "multigrid.cu", "io.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # no FMA contraction: bit parity with the reference
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math",
    "-Xptxas", "-v" if os.environ.get("O3D_PTXAS_V") else "-warn-spills",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "o3d_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc:
                    print(" ".join(cmd))
                    print(out)
                elif out.strip():
                    print(out.strip())
                if rc:
                    raise RuntimeError("nvcc failed for %s" % cmd[-3])
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            print(r.stdout)
            raise RuntimeError("link failed")
    build_mainloop(force)
    return LIB


MAINLOOP_SRC = os.path.join(HERE, "host", "o3d_mainloop.cpp")
MAINLOOP_EXE = os.path.join(OUT_DIR, "o3d_mainloop")


def build_mainloop(force=False):
    """the driver-shaped harness over include/o3d_b200.hpp (host/o3d_mainloop.cpp): plain g++,
    linked against the in-tree library with an $ORIGIN rpath so that it travels with it"""
    gxx = shutil.which("g++")
    if not gxx:
        return None
    hpp = os.path.join(HERE, "..", "include", "o3d_b200.hpp")
    if force or _stale(MAINLOOP_EXE, [MAINLOOP_SRC, hpp, LIB]):
        cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-o", MAINLOOP_EXE,
               MAINLOOP_SRC, "-L" + OUT_DIR, "-lo3d_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            print(r.stdout)
            raise RuntimeError("g++ failed for o3d_mainloop.cpp")
    return MAINLOOP_EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
