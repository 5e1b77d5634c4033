"""Out-of-bounds stores: every kernel family once (profiles/sanitize_case.py) with KMER_B200_GUARD=1 -- canary zones
around every device allocation of the library, verified on the device when the allocation is freed. This stands in for
compute-sanitizer's memcheck, which the GPU pool does not allow."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_store_lands_outside_a_device_buffer():
    env = dict(os.environ, KMER_B200_GUARD="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sanitize_case.py")], capture_output=True, text=True,
                         timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sanitize case done" in out.stdout
    assert "guard violations 0 selftest ok True" in out.stdout, out.stdout[-2000:]
