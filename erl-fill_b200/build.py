"""In-tree build of liberlfill.so (nvcc, sm_100a only).  Used by __graft_entry__.build().

    python erl-fill_b200/build.py [--force] [--verbose]

Every .cu under csrc/ is compiled to an object with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --std=c++17 -Xcompiler -fPIC
and linked into erl-fill_b200/lib/liberlfill.so (cudart linked statically, so the library only
needs the driver at run time).  env_step.cu additionally gets -fmad=false (bit-exact binary64
simulator, see its header).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
SO = os.path.join(LIBDIR, "liberlfill.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "--std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC, "--expt-relaxed-constexpr"]
PER_FILE = {"env_step.cu": ["-fmad=false"]}
if os.environ.get("ERLFILL_TC_DEFS"):       # development: e.g. ERLFILL_TC_DEFS="-DERLFILL_TC_TILES=2" (defaults live in the source)
    PER_FILE["emotion_tf_tc.cu"] = os.environ["ERLFILL_TC_DEFS"].split()


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest(paths, extra):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "erlfill.h"))
    objs, rebuilt = [], False
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        flags = ARCH + COMMON + PER_FILE.get(src, [])
        stamp = obj + ".sha"
        dig = _digest([path] + headers, flags)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            objs.append(obj)
            continue
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        with open(stamp, "w") as f:
            f.write(dig)
        objs.append(obj)
        rebuilt = True
    if rebuilt or force or not os.path.exists(SO):
        cmd = [nvcc] + ARCH + ["-shared", "-o", SO] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
