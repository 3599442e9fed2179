"""Stage the UNMODIFIED reference sources of the hot path for the GPU box.

    python oracle/make_ref.py

The reference is pure Python and lives in /root/reference, which does not exist on
the GPU box.  This script copies the four files the path consists of
(subpixal/{utils,centroid,cc,cutout}.py) byte for byte into ``oracle/_ref/subpixal/``
- a git-IGNORED directory (it never enters the history) that is NOT gpurun-ignored, so
it travels to the GPU box like a built ``.so`` does.  ``oracle/refload.py`` falls back
to it when /root/reference is absent; ``bench.py`` then times the live reference
(``cpu_baseline.kind = "reference"``) instead of the oracle port.  Test
infrastructure only: nothing under ``subpixal_b200/`` imports it.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SUBPIXAL_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("utils.py", "centroid.py", "cc.py", "cutout.py")


def stage(verbose=False):
    """Copy the files if the reference tree is mounted; returns True when oracle/_ref is complete."""
    src_pkg = os.path.join(SRC, "subpixal")
    if os.path.isfile(os.path.join(src_pkg, "cc.py")):
        os.makedirs(os.path.join(DST, "subpixal"), exist_ok=True)
        sums = []
        for f in FILES:
            shutil.copyfile(os.path.join(src_pkg, f), os.path.join(DST, "subpixal", f))
            with open(os.path.join(src_pkg, f), "rb") as fh:
                sums.append("%s  %s" % (hashlib.sha256(fh.read()).hexdigest(), f))
        with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
            fh.write("\n".join(sums) + "\n")
        if verbose:
            print("\n".join(sums))
    return all(os.path.isfile(os.path.join(DST, "subpixal", f)) for f in FILES)


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("oracle/_ref %s" % ("ready" if ok else "NOT available (reference tree not mounted)"))
    sys.exit(0 if ok else 1)
