"""Build libfocalsv_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

The shared object lands next to this file (focalsv_b200/libfocalsv_cuda.so): it is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfocalsv_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "--use_fast_math", "-Xptxas", "-v",
         "-I", os.path.join(HERE, "..", "include")]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, "..", "include", "focalsv_cuda.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB, os.path.join(CSRC, "fsv_capi.cu")]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as fh:
        fh.write(r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed (see %s)" % log)
    if verbose:
        print(r.stdout)
    return LIB


def build_variant(out, defs):
    """Experiment builds: same sources with extra -D flags into another file (scripts/ only)."""
    cmd = [NVCC] + FLAGS + list(defs) + ["-o", out, os.path.join(CSRC, "fsv_capi.cu")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
