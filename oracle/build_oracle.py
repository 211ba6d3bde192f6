"""oracle/build_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Builds the two CPU checkers:

* ``oracle/_build/liboracle.so``  -- the C restatement (libdistance_oracle.c);
  always buildable (gcc only).
* ``oracle/_ref/libref.so``       -- the UNMODIFIED reference libdistance headers
  and kmedoids.cc, compiled from where they lie under /root/reference through
  ref_shim.cc.  Only possible in the build container (the GPU box has no
  /root/reference); the .so is git-ignored but travels with gpurun.

Called by ``__graft_entry__.build()``; may also be run by hand:
    python oracle/build_oracle.py
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MSMB_REFERENCE_ROOT", "/root/reference")


def _newer(target, *sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_restatement(force=False):
    src = os.path.join(HERE, "libdistance_oracle.c")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liboracle.so")
    if force or _newer(out, src):
        # -O2 without -ffast-math: keep IEEE evaluation order (parity oracle).
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-fno-fast-math",
               "-ffp-contract=off", src, "-o", out, "-lm"]
        subprocess.check_call(cmd)
    return out


def build_reference(force=False):
    """Compile the real reference sources (if present).  Returns path or None."""
    libdist = os.path.join(REFERENCE_ROOT, "msmbuilder", "libdistance", "src")
    cluster = os.path.join(REFERENCE_ROOT, "msmbuilder", "cluster", "src")
    out_dir = os.path.join(HERE, "_ref")
    out = os.path.join(out_dir, "libref.so")
    if not (os.path.isdir(libdist) and os.path.isdir(cluster)):
        return out if os.path.exists(out) else None
    import numpy
    os.makedirs(out_dir, exist_ok=True)
    shim = os.path.join(HERE, "ref_shim.cc")
    if force or _newer(out, shim):
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++11", "-w",
               "-ffp-contract=off",
               "-DPyInt_AsLong=PyLong_AsLong",
               "-I", libdist, "-I", cluster,
               "-I", numpy.get_include(),
               "-I", sysconfig.get_paths()["include"],
               shim, "-o", out]
        subprocess.check_call(cmd)
    return out


def main():
    a = build_restatement(force="--force" in sys.argv)
    b = build_reference(force="--force" in sys.argv)
    print("restatement:", a)
    print("reference  :", b)


if __name__ == "__main__":
    main()
