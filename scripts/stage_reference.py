"""Stage the UNMODIFIED reference package into the git-ignored ``baseline/_ref/`` so that
``bench.py --impl reference`` can time ``pg.Lagrange1().assemble_stiff_matrix`` itself on the GPU
box's host cores.  Build container only (needs /root/reference).

    python scripts/stage_reference.py

First choice is the offline pip install the bench contract names; the reference declares porepy
(a git dependency, absent) so it runs with ``--no-deps``; if pip cannot build the wheel offline the
package directory is copied as it is.  Nothing under baseline/_ref is tracked by git; gpurun ships it.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DEST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(REF, "src", "pygeon")):
        raise RuntimeError("reference source not present")
    how = None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "docs", "tutorials", "tests"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode == 0 and os.path.isdir(os.path.join(DEST, "pygeon")):
            how = "pip install --no-index --no-build-isolation --no-deps --target baseline/_ref"
        else:
            if verbose:
                print(r.stdout[-2000:], file=sys.stderr)
            shutil.copytree(os.path.join(REF, "src", "pygeon"), os.path.join(DEST, "pygeon"))
            how = "copied src/pygeon (pip could not build the wheel offline)"
    with open(os.path.join(DEST, "STAGED.txt"), "w") as f:
        f.write(how + "\n")
    if verbose:
        print(f"baseline/_ref staged: {how}")
    return how


if __name__ == "__main__":
    stage()
