"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY).

  oracle/_ref/libfsv_oracle.so  <- oracle/ksw2_oracle.c  (the plain-C restatement)
  oracle/_ref/libksw2_ref.so    <- the reference's OWN, unmodified
        /root/reference/software/hifiasm-0.16.1/ksw2_extz2_sse.c (+ ksw2.h),
        compiled where it lies together with oracle/ref_shim.c.  Built only when
        /root/reference exists (this container); the GPU box uses the prebuilt
        .so that travels with the snapshot.  No reference source is copied.

Both outputs are git-ignored (oracle/_ref/) and NOT gpurun-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_DIR = "/root/reference/software/hifiasm-0.16.1"
CFLAGS = ["-O3", "-msse4.1", "-std=gnu11", "-fPIC", "-shared", "-fno-strict-aliasing", "-pthread"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    built = {}
    port = os.path.join(OUT, "libfsv_oracle.so")
    srcs = [os.path.join(HERE, "ksw2_oracle.c"), os.path.join(HERE, "ksw2_oracle.h"), os.path.join(HERE, "batch_pool.h"),
            os.path.join(HERE, "..", "include", "focalsv_cuda.h")]
    if _stale(port, srcs):
        cmd = ["gcc"] + CFLAGS + ["-o", port, srcs[0]]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    built["port"] = port
    ref = os.path.join(OUT, "libksw2_ref.so")
    ref_src = os.path.join(REF_DIR, "ksw2_extz2_sse.c")
    if os.path.exists(ref_src):
        shim = os.path.join(HERE, "ref_shim.c")
        if _stale(ref, [shim, ref_src, os.path.join(REF_DIR, "ksw2.h"), os.path.join(HERE, "batch_pool.h")]):
            cmd = ["gcc"] + CFLAGS + ["-I", REF_DIR, "-o", ref, shim, ref_src]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if os.path.exists(ref):
        built["reference"] = ref
    return built


if __name__ == "__main__":
    print(build(verbose=True))
