#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference's Python package, so that the reference
ITSELF (not the oracle port) can run on the GPU box — as the CPU arm of ``bench.py --impl reference`` and as the fp32
eager same-device comparison of the GPU tests.  ``/root/reference`` does not exist on the GPU box; ``oracle/_ref/`` is
listed in .gitignore (never in history) but not in .gpurunignore, so it travels with the snapshot like a built .so.

The reference is pure Python (no build system, no native code): "building" it is copying ``expertsim/`` and ``cli.py``
where they lie.  Nothing is edited; ``oracle/ref_shim.py`` supplies the import stubs at run time.  Called by
``__graft_entry__.build()`` whenever ``/root/reference`` is present.
"""
import hashlib
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def tree_digest(root):
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py") or f.endswith(".yaml"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def make_ref(verbose=True):
    """-> True if oracle/_ref is in place (fresh copy, or already identical), False if there is no reference here."""
    if not os.path.isdir(os.path.join(SRC, "expertsim")):
        if verbose:
            print(f"[make_ref] {SRC} not present: leaving {DST} as it is")
        return os.path.isdir(os.path.join(DST, "expertsim"))
    want = tree_digest(os.path.join(SRC, "expertsim"))
    stamp = os.path.join(DST, "DIGEST")
    if os.path.exists(stamp) and open(stamp).read().strip() == want:
        return True
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "expertsim"), os.path.join(DST, "expertsim"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy(os.path.join(SRC, "cli.py"), os.path.join(DST, "cli.py"))
    open(stamp, "w").write(want + "\n")
    if verbose:
        print(f"[make_ref] copied {SRC}/expertsim -> {DST} (digest {want})")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
