"""Builds libcrlk.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels
with the repository snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcrlk.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--expt-extended-lambda", "--expt-relaxed-constexpr",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
# (source, extra flags).  The float64 planner kernels must not contract a*b+c.
EXTRA_DEFS = ["-D" + d for d in os.environ.get("CRLK_DEFS", "").split() if d]
UNITS = [("crlk_kernels.cu", ["--fmad=false"] + EXTRA_DEFS),
         ("crlk_planner.cu", ["--fmad=false"]),
         ("crlk_nets.cu", ["--fmad=false"] + EXTRA_DEFS),   # shares the float64 helpers of crlk_device.cuh; float32 FMAs are explicit
         ("crlk_mlp_tc.cu", EXTRA_DEFS)]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "crlk.h"))
    return out


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [_nvcc()] + ARCH + ["-shared", "-o", SO] + objs + ["-lcuda"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
