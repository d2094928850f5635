"""Build the in-tree CUDA library (anyfeature-vslam_b200/libafv_b200.so) with nvcc for sm_100a.

Pure CUDA-runtime shared library with a C ABI (include/afv.h): no torch / pybind types, loadable with ctypes
or linked from C++.  nvcc cross-compiles without a GPU, so this also runs in the CPU build container.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libafv_b200.so")
# (source, extra flags).  The extraction kernels must round like the pinned CPU path: no implicit FMA.
SOURCES = [
    ("afv_orb.cu", ["--fmad=false"]),
    ("afv_match.cu", ["--fmad=false"]),
    ("afv_sift.cu", ["--fmad=false"]),
    ("afv_akaze.cu", ["--fmad=false"]),
    ("afv_brisk.cu", ["--fmad=false"]),
    ("afv_orbslam2.cu", ["--fmad=false"]),
    ("afv_capi.cu", ["--fmad=false"]),
]
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "--expt-relaxed-constexpr"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return p


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "afv.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = nvcc_path()
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src, extra in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, "-c", sp, "-o", obj] + COMMON + extra + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        if verbose or "warning" in out:
            sys.stderr.write(out)
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
