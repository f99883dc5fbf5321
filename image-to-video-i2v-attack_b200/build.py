"""Compile the CUDA sources under csrc/ into libi2v_b200.so, in-tree, for sm_100a only.

nvcc cross-compiles without a GPU; the built library is git-ignored but travels to the GPU box with
the gpurun snapshot.  `python -m i2v_b200.build` or `__graft_entry__.build()`.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libi2v_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]
# update.cu carries the bit-exact rounding contract: no implicit FMA contraction at all.
PER_FILE = {"update.cu": ["-fmad=false"]}


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; i2v_b200 has no prebuilt or CPU fallback")


def _digest(paths, extra):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Build (or reuse an up-to-date) libi2v_b200.so; returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "i2v_b200.h"))
    headers.append(os.path.join(HERE, "..", "include", "i2v_b200_debug.h"))
    stamp = os.path.join(OBJ, "stamp.txt")
    want = _digest([os.path.join(CSRC, s) for s in sources] + headers, (ARCH, COMMON, PER_FILE))
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read().strip() == want:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in sources:
        o = os.path.join(OBJ, s[:-3] + ".o")
        cmd = [nvcc] + ARCH + COMMON + PER_FILE.get(s, []) + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out.decode(errors="replace")))
        if verbose and out:
            print(out.decode(errors="replace"), file=sys.stderr)
    # cudart is linked statically (nvcc default); the driver API (TMA descriptors) is resolved at run
    # time through cudaGetDriverEntryPoint, so the library loads on a CPU-only box for symbol checks.
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", LIB] + objs)
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
