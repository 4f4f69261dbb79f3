"""Cold-launch regression guard (was scripts/gpu_res.sh): the CTA-pair conv kernel once returned stale residual-tile
rows in the FIRST tile a CTA reads, on cold launches only (a WAR race between a TMA refill and outstanding
shared-memory loads).  A warm pytest process cannot see that class of bug, so the residual /
CTA-pair / staged-store conv tests are re-run here in four FRESH Python processes, each of which launches the kernels
cold (new context, cold instruction cache, empty L2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = "staged_tma_store or cta_pair or kitti_level or residual_modes"


@pytest.mark.parametrize("attempt", range(4))
def test_conv_kernels_cold_process(attempt):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_conv3d.py"), "-q", "-x", "-m", "gpu",
                        "-k", SELECT, "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], tail
