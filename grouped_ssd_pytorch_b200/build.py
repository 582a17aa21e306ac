"""Build libgssd_b200.so in-tree with nvcc for sm_100a (no torch headers: the boundary is a C ABI).

    python -m grouped_ssd_pytorch_b200.build [--force]

The .so lands next to this file so that it travels with the source tree; it is git-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgssd_b200.so")
SOURCES = ["abi.cu", "boxes.cu", "match.cu", "loss.cu", "fused.cu", "detect.cu", "evalap.cu", "gconv.cu", "gconv_bwd.cu", "dcn.cu", "attn.cu", "bnrelu.cu", "pipe.cu"]
HEADERS = ["common.cuh", "select.cuh", "tc.cuh", "tmap.cuh", os.path.join("..", "..", "include", "gssd.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact float contract: no FMA contraction, IEEE div/sqrt, no flush-to-zero
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xcudafe", "--diag_suppress=177",
    "-cudart", "static",
]
# gconv*.cu / dcn.cu (bf16 tensor-core path, 1e-2 tolerance) and attn.cu (fp32 dot products, 1e-5) carry no bit-exact contract:
# FMA contraction allowed
PER_SOURCE_FLAGS = {"gconv.cu": ["-fmad=true"], "gconv_bwd.cu": ["-fmad=true"], "dcn.cu": ["-fmad=true"], "attn.cu": ["-fmad=true"], "bnrelu.cu": ["-fmad=true"]}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, phase_timing=False):
    """phase_timing=True builds libgssd_b200_dbg.so (clock64 stamps per kernel phase, development only)."""
    lib = LIB.replace(".so", "_dbg.so") if phase_timing else LIB
    if not force and not phase_timing and not stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    extra = ["-DGSSD_PHASE_TIMING"] if phase_timing else []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", "_dbg.o" if phase_timing else ".o"))
        flags = [f for f in NVCC_FLAGS if not (s in PER_SOURCE_FLAGS and f == "-fmad=false")] + PER_SOURCE_FLAGS.get(s, [])
        cmd = [_nvcc()] + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libgssd_b200.so")
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", lib] + objs
    subprocess.check_call(link)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phase_timing="--phase-timing" in sys.argv))
