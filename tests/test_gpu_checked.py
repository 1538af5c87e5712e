"""The GPU tests once more against the CHECKED build of the library (make -C mouse_b200/csrc checked:
-DMOUSE_CHECKED compiles device-side bounds assertions into the sweeps, the cell build and the dump parser -
compute-sanitizer is not available on the GPU pool, this is the bounds evidence that is).  A fired assertion aborts
the child process and prints the failing expression."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "mouse_b200", "libmouse_b200_checked.so")


def test_gpu_suites_pass_against_the_checked_library():
    if not os.path.exists(CHECKED):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mouse_b200", "csrc"), "checked"])
    assert b"dst + cnt <= LZ" in open(CHECKED, "rb").read()           # the assertions are compiled in ...
    assert b"dst + cnt <= LZ" not in open(os.path.join(ROOT, "mouse_b200", "libmouse_b200.so"), "rb").read()   # ... here only
    env = dict(os.environ, MOUSE_B200_LIB=CHECKED)
    files = ["tests/test_gpu_stress.py", "tests/test_gpu_parity.py", "tests/test_gpu_ingest.py", "tests/test_gpu_fullsize.py",
             "tests/test_gpu_triclinic.py"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + files,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert "Assertion" not in r.stdout + r.stderr, tail
    assert r.returncode == 0, tail
    assert " passed" in r.stdout
