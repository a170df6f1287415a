"""Build lib/libgpvb200.so from csrc/*.cu with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libgpvb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]


def source_sha16():
    """Identity of the library's CODE: sha256 over csrc/*, include/*.h and the nvcc flags.  (The .so's own bytes differ from build to
    build by ~90 bytes of embedded ids, so a file hash cannot tie a profile to a build.)"""
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    files = sorted(glob.glob(os.path.join(CSRC, "*"))) + sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(s, o) or hdr_time > os.path.getmtime(o):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {s}:\n{out}\n")
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("gpvb200 build failed")
    if force or procs or not os.path.exists(SO):
        cmd = [NVCC, "-arch=sm_100a", "-shared", "-o", SO] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
