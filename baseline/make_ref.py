#!/usr/bin/env python3
"""baseline/make_ref.py -- MEASUREMENT / TEST INFRASTRUCTURE.

Copies the UNMODIFIED reference package (src/specimux), its own tests and its test data from /root/reference
into the git-ignored baseline/_ref/ so that they travel to the GPU box (where /root/reference does not exist):

  * bench.py --impl reference runs that copy's own CLI (`python -m specimux.cli ... -F -t <cores>`) over the
    stand-ins of oracle/standins (edlib -> the C restatement, Bio, pybloomfilter; none of the three is
    installable here or on the box) -- the reference arm, kind "reference+shim";
  * tests/test_reference_suite.py runs the reference's own test files against the `specimux` alias package.

Nothing under baseline/_ref is imported by the product, and nothing of it is committed.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")


def make(force=False):
    if not os.path.isdir(os.path.join(REF, "src", "specimux")):
        return os.path.isdir(os.path.join(DST, "src", "specimux"))
    stamp = os.path.join(DST, ".copied")
    if os.path.exists(stamp) and not force:
        return True
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, "src"), os.path.join(DST, "src"), ignore=shutil.ignore_patterns("__pycache__", "*.egg-info"))
    shutil.copytree(os.path.join(REF, "tests"), os.path.join(DST, "tests"), ignore=shutil.ignore_patterns("__pycache__"))
    for f in ("pyproject.toml", "LICENSE"):
        if os.path.exists(os.path.join(REF, f)):
            shutil.copy(os.path.join(REF, f), os.path.join(DST, f))
    open(stamp, "w").write("copied from %s\n" % REF)
    return True


if __name__ == "__main__":
    ok = make(force="--force" in sys.argv)
    print("baseline/_ref %s" % ("ready" if ok else "unavailable (no /root/reference here)"))
