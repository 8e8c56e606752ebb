"""Randomised parity campaign (tools/parity_campaign.py): random scenes of 1..300
boxes and spheres (shared-memory scan and LBVH), random camera poses and fovs,
scales, column counts, pass indices, all three kernels -- CUDA frame and ray count
vs the oracle, bit for bit.  9000 cases / 36 M rays were run clean in round 1;
the suite keeps 500."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_random_scenes_cameras_kernels_bit_exact():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_campaign.py"), "--cases", "500", "--seed", "11"],
                         capture_output=True, text=True, timeout=900)
    tail = [l for l in out.stdout.splitlines() if "cases" in l or "MISMATCH" in l]
    assert out.returncode == 0, "\n".join(tail[-10:]) + out.stderr[-2000:]
    assert tail and "500 cases, 0 mismatches" in tail[-1]
