"""Build libstereosvo_b200.so (CUDA kernels + C-ABI + host facade) in-tree for sm_100a.

    python -m stereo_svo_slam_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libstereosvo_b200.so")
SOURCES = ["rectify.cu", "pyramid.cu", "align.cu", "klt.cu", "refine.cu", "stereo.cu", "detect.cu", "context.cu", "slam_host.cpp"]
HEADERS = ["common.cuh", "kernels.cuh", os.path.join("..", "..", "include", "svo_cuda.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # the reference's float arithmetic is un-fused scalar C++; keep per-keypoint values bit-comparable
              "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, logs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append(f"== {s}\n{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {s}")
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    if logs:
        with open(os.path.join(objdir, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
