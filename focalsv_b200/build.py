"""Build libfocalsv_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

The shared object lands next to this file (focalsv_b200/libfocalsv_cuda.so): it is
git-ignored but travels to the GPU box with the gpurun snapshot.  The DPX fill kernel is
compiled as eight translation units (dual x traceback/approx mode), in parallel, then linked with the
C ABI (fsv_capi.cu).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libfocalsv_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
          "-Xcompiler", "-fPIC,-O3,-Wall", "--use_fast_math", "-Xptxas", "-v",
          "-I", os.path.join(HERE, "..", "include")]
VARIANTS = [(d, t) for d in (0, 1) for t in (0, 1, 2, 3)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, "..", "include", "focalsv_cuda.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(s) > t for s in deps)


def _units(defs):
    units = [("capi", os.path.join(CSRC, "fsv_capi.cu"), []), ("chain", os.path.join(CSRC, "fsv_chain.cu"), [])]
    for d, t in VARIANTS:
        units.append(("dpx_%d%d" % (d, t), os.path.join(CSRC, "fsv_dpx_variant.cu"),
                      ["-DFSV_VARIANT_DUAL=%d" % d, "-DFSV_VARIANT_TBM=%d" % t]))
    return [(n, s, f + list(defs)) for n, s, f in units]


def _deps(src):
    """Files a translation unit is rebuilt for: itself, the headers it includes (recursively, csrc/ and include/ only)."""
    seen, todo = set(), [src]
    while todo:
        f = todo.pop()
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        for ln in open(f, errors="replace"):
            ln = ln.strip()
            if ln.startswith("#include \""):
                inc = ln.split('"')[1]
                todo.append(os.path.normpath(os.path.join(os.path.dirname(f), inc)))
    return seen | {os.path.abspath(__file__)}


def _compile(objdir, name, src, flags):
    obj = os.path.join(objdir, name + ".o")
    stamp = obj + ".flags"
    want = " ".join(CFLAGS + flags)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want and \
            all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in _deps(src)):
        return name, obj, 0, "(up to date)\n"
    r = subprocess.run([NVCC] + CFLAGS + flags + ["-c", "-o", obj, src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode == 0:
        with open(stamp, "w") as fh:
            fh.write(want)
    return name, obj, r.returncode, r.stdout


def _build(out, defs, objdir, log):
    os.makedirs(objdir, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        res = list(ex.map(lambda u: _compile(objdir, *u), _units(defs)))
    text = "".join("==== %s\n%s" % (n, o) for n, _, _, o in res)
    if log:
        with open(log, "w") as fh:
            fh.write(text)
    if any(rc for _, _, rc, _ in res):
        sys.stderr.write(text)
        raise RuntimeError("nvcc failed" + (" (see %s)" % log if log else ""))
    r = subprocess.run([NVCC, "-shared", "-o", out] + [o for _, o, _, _ in res], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return out


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    _build(LIB, [], OBJ, os.path.join(HERE, "build.log"))
    if verbose:
        print(open(os.path.join(HERE, "build.log")).read())
    return LIB


def build_variant(out, defs):
    """Experiment builds: same sources with extra -D flags into another file (scripts/ only)."""
    return _build(out, defs, OBJ + "_" + os.path.basename(out), None)


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
