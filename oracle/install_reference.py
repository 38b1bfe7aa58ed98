#!/usr/bin/env python
"""Put the reference package where bench.py --impl reference can import it on the GPU box.

    python oracle/install_reference.py        ->  baseline/_ref/numbacs/   (git-ignored)

The reference (alb3rtjarvis/numbacs v0.1.2) cannot be pip-installed here: its pyproject requires
Python < 3.12 and the hatchling backend, and its dependencies numbalsoda / interpolation /
contourpy are not in the offline wheelhouse.  It is pure Python, so "installing" it is copying its
package directory; the two missing third-party leaves come from oracle/shims (see README there).
baseline/_ref is listed in .gitignore (reference sources never enter the history) but not in
.gpurunignore, so the copy travels to the GPU box with the working tree.  Called by
__graft_entry__.build() when /root/reference is present."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/numbacs"
DST = os.path.join(ROOT, "baseline", "_ref", "numbacs")


def install(force=False):
    if not os.path.isdir(SRC):
        return None
    if os.path.isdir(DST) and not force:
        return DST
    if os.path.isdir(DST):
        for root, dirs, files in os.walk(DST):
            for name in dirs:
                os.chmod(os.path.join(root, name), 0o755)
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__"))
    for root, dirs, files in os.walk(os.path.dirname(DST)):     # the source tree is read-only; the copy is not
        for name in dirs + files:
            pth = os.path.join(root, name)
            os.chmod(pth, os.stat(pth).st_mode | 0o200)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
