"""In-tree build of liblidarreg.so (hand-written sm_100a CUDA behind a C ABI).

`python -m lidarregistration_b200.build` or `__graft_entry__.build()`.
nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the
GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# LIDARREG_SO: load (and build into) another file than the product library -- A/B builds of a kernel (tools/build_variants.py)
SO = os.environ.get("LIDARREG_SO") or os.path.join(CSRC, "liblidarreg.so")
SOURCES = ["lr_core.cu", "lr_prof.cu", "lr_ransac.cu", "lr_match.cu", "lr_match_tc.cu", "lr_gpf.cu"]
HEADERS = ["lr_common.cuh", "lr_match_tc.cuh", "lr_ransac_gc.cuh", "lr_score_tc.cuh", "lr_icp.cuh", "lr_icp_api.cuh", os.path.join("..", "..", "include", "lidarreg.h")]

# -fmad=false: the canonical fp64/fp32 arithmetic must not be contracted behind
# our back; every FMA the kernels want is written as fmaf()/fma() explicitly.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblidarreg.so cannot be built (no CPU fallback exists)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    # LIDARREG_NVCC_FLAGS: extra -D switches for A/B builds of a kernel (tools/); empty in the product build
    extra = os.environ.get("LIDARREG_NVCC_FLAGS", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + sources()
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc; let nvcc pick its default host compiler
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building liblidarreg.so")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
