#!/usr/bin/env python
"""Run the reference's OWN test files, unchanged, against its OWN source over the shims.

    python oracle/run_reference_tests.py [pytest args]

PYTHONPATH = oracle/shims : <reference>/src, where <reference> is /root/reference (this container)
or baseline/_ref (a copy made by oracle/install_reference.py; travels to the GPU box).  The test
files come from /root/reference/tests when present (they are read, not copied).  This is what
accepts the shims and what pins the oracle's integrator a second,
independent time: the same C DOP853 reproduces fm.npy / fm_n.npy / fm_aux.npy / ftle.npy / lavd.npy
when driven by the reference's own Python."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_src():
    for cand in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "numbacs")):
            return cand
    return None


def env_with_shims():
    src = reference_src()
    if src is None:
        raise RuntimeError("no reference source: neither /root/reference/src nor baseline/_ref/numbacs exists")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), src, env.get("PYTHONPATH", "")])
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_ref")
    return env


def main(argv):
    tests = "/root/reference/tests"
    if not os.path.isdir(tests):
        print("the reference's tests/ are not available here (they live in /root/reference only)")
        return 0
    files = [os.path.join(tests, f) for f in ("test_integration.py", "test_flows.py", "test_diagnostics.py")]
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp"] + files + argv
    return subprocess.call(cmd, env=env_with_shims(), cwd="/tmp")


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
