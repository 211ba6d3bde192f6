"""Build libmsmb200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m msmbuilder_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libmsmb200.so")
SOURCES = ["lib.cu", "dist_kernels.cu", "kcenters_lookahead.cu", "tica_simt.cu", "tica_umma.cu", "assign_umma.cu", "rmsd.cu",
           "scan_kernels.cu", "kmedoids_host.cpp"]
HEADERS = ["common.cuh", "tica_umma_v2.cuh", os.path.join("..", "..", "include", "msmb200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-cudart", "static"]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, h) for h in HEADERS]
    extra = os.path.splitext(src)[0] + ".cuh"
    if os.path.exists(extra):
        deps.append(extra)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            if s.endswith(".cpp"):
                cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("[%s]\n%s\n" % (s, out.decode(errors="replace")))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(OUT):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
               "-o", OUT] + objs + ["-lpthread", "-ldl", "-lrt"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
