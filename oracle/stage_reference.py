#!/usr/bin/env python
"""Stage the UNMODIFIED reference (PyREMOT) under oracle/_ref/ so that it can be timed on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing under rmt_app_b200/ imports this file or anything it produces; only
tests/, __graft_entry__ (build/smoke) and bench.py's CPU legs do.

The reference is pure Python (numpy + scipy; matplotlib only for figures): there is nothing to compile, "building" it
means copying the package tree it is imported from.  /root/reference does not exist on the GPU box, so build() runs
this script in the build container: it copies the files of the path (SURVEY.md 8(c) list plus the modules the package
imports eagerly at `import PyREMOT` — rmtCore.py:8-19 pulls in every model module) byte for byte into oracle/_ref/, which
is git-ignored (the sources never enter the history) but not gpurun-ignored (it travels to the box like the built .so).
A file manifest with SHA-256 sums of source and copy is written beside it (oracle/_ref/MANIFEST.json).

usage: python oracle/stage_reference.py [--reference /root/reference] [--force]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# package sub-trees needed to import PyREMOT and run rmtExe on models N1 / N2 (examples, tests and notebooks are not)
SUBTREES = ("core", "data", "docs", "examples", "library", "solvers")   # examples: imported eagerly by docs/rmtCore.py:18
TOP_FILES = ("__init__.py", "rmt.py")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(reference="/root/reference", force=False, quiet=False):
    """Returns the staged package root (oracle/_ref) or None when the reference tree is absent (GPU box: the
    directory staged in the build container is used as it is)."""
    src_pkg = os.path.join(reference, "PyREMOT")
    man_path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isdir(src_pkg):
        return DEST if os.path.exists(man_path) else None
    files = [os.path.join(src_pkg, f) for f in TOP_FILES]
    for sub in SUBTREES:
        for root, dirs, names in os.walk(os.path.join(src_pkg, sub)):
            dirs[:] = sorted(d for d in dirs if d != "__pycache__")
            files += [os.path.join(root, n) for n in sorted(names) if n.endswith((".py", ".json", ".txt"))]
    if not force and os.path.exists(man_path):
        try:
            old = json.load(open(man_path))
            if all(old["files"].get(os.path.relpath(f, reference)) == _sha(f) for f in files) and \
                    len(old["files"]) == len(files):
                return DEST
        except Exception:
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, reference)
        out = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
        assert _sha(out) == _sha(f)
        manifest[rel] = _sha(f)
    with open(man_path, "w") as fh:
        json.dump({"reference": reference, "files": manifest,
                   "note": "byte-for-byte copy of the reference files of the N1/N2 path; unmodified"}, fh, indent=1)
    if not quiet:
        print("[stage_reference] %d files -> %s" % (len(files), DEST))
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("RMT_REFERENCE", "/root/reference"))
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    r = stage(a.reference, a.force)
    print(r or "reference tree not found and nothing staged")
    sys.exit(0 if r else 1)
