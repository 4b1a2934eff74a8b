"""Build libcproc_cuda.so (sm_100a only) and the C drop-in shims, in-tree.

    python synth_tools_b200/build.py [--force] [--ptxas-v]   (run by path: importing the package dlopens the library)

nvcc cross-compiles without a GPU.  The built .so files stay in the package
directory (git-ignored) so they travel with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libcproc_cuda.so")
DROPIN = os.path.join(HERE, "libcproc_dropin.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=False, ptxas_v=False):
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(ROOT, "include", "cproc_cuda.h")]
    if force or _newer(LIB, deps):
        objdir = os.path.join(HERE, "build")
        os.makedirs(objdir, exist_ok=True)
        objs = []
        procs = []
        for s in srcs:
            o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
            objs.append(o)
            if not force and not _newer(o, [s] + deps[len(srcs):]):
                continue
            cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O2", *os.environ.get("CPROC_NVCC_DEFS", "").split(),
                   "-I", os.path.join(ROOT, "include"), "-c", s, "-o", o]
            if ptxas_v:
                cmd += ["-Xptxas", "-v"]
            if verbose:
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd)))
        for s, p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed on %s" % s)
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    # m_pd.h present (PD_INCLUDE=<dir>, or a system Pd): the drop-in library also exports square_grain_proc under the
    # reference's own name and prototype (linux/synth_tools.c:85-86); host/dropin.c, CPROC_HAVE_PD
    pd_flags = []
    for d in [os.environ.get("PD_INCLUDE"), "/usr/include", "/usr/include/pd", "/usr/local/include", "/usr/local/include/pd"]:
        if d and os.path.exists(os.path.join(d, "m_pd.h")):
            pd_flags = ["-DCPROC_HAVE_PD", "-I", d]
            break
    host_srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".c")) if os.path.isdir(HOST) else []
    if host_srcs and (force or _newer(DROPIN, host_srcs + [LIB])):
        cmd = ["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"), *pd_flags,
               *host_srcs, "-o", DROPIN, "-L", HERE, "-lcproc_cuda", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas-v" in sys.argv)
    print("built", LIB)
