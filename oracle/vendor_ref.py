"""Vendors the UNMODIFIED reference tree into the git-ignored `baseline/_ref/erlfill_reference/` (test infrastructure).

    python oracle/vendor_ref.py            # copy /root/reference -> baseline/_ref/erlfill_reference + MANIFEST.json

Why: the reference is pure Python, so there is nothing to compile into `oracle/_ref`; but `bench.py --impl reference` and the
`cpu_baseline` leg must time the reference's OWN implementation on the GPU box's host cores, where `/root/reference` does not
exist.  `baseline/_ref/` is git-ignored (history stays free of reference sources) and NOT gpurun-ignored, so the byte-for-byte
copy travels with the snapshot.  `__graft_entry__.build()` calls `vendor()` whenever `/root/reference` is present.
Nothing in the product package reads this directory; only `bench.py`'s CPU legs and the GPU test that runs the reference's
`train_ddpg` against the drop-ins do (through `oracle/ref_shim.py`, which stubs `pymodbus`).
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref", "erlfill_reference")
KEEP = (".py", ".txt", ".md")


def _files(base):
    out = []
    for d, dirs, files in os.walk(base):
        dirs[:] = sorted(x for x in dirs if x not in (".git", "__pycache__"))
        for f in sorted(files):
            if f.endswith(KEEP):
                out.append(os.path.relpath(os.path.join(d, f), base))
    return out


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def vendor(src=SRC, dst=DST):
    """Byte-for-byte copy; returns the manifest {relative path: sha256}.  No-op (returns None) without the source tree."""
    if not os.path.isdir(src):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    manifest = {}
    for rel in _files(src):
        to = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(to), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), to)
        manifest[rel] = _sha(to)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dst=DST):
    """True when every vendored file still has the recorded hash (i.e. the copy is unmodified)."""
    mf = os.path.join(dst, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    files = json.load(open(mf))["files"]
    return all(os.path.exists(os.path.join(dst, r)) and _sha(os.path.join(dst, r)) == h for r, h in files.items())


if __name__ == "__main__":
    m = vendor()
    print("no reference tree at %s" % SRC if m is None else "vendored %d files into %s" % (len(m), DST))
    sys.exit(0)
