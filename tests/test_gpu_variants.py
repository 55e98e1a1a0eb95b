"""GPU: the switchable kernel variants must give the same bits as the default forms.

``RDP_PFN_ROWS=1`` selects the thread = row forward (``pfn_rows_kernel``) and ``RDP_BWD_STREAM=1`` the pillar-streaming
backward (``pfn_bwd_stream_kernel``).  librdp reads the switches once per process, so the parity suite is re-run in a
child process with both set: every golden, the two-frame LiDAR case and the stress cloud go through the variant kernels
and are held to the same bars (bit-exact integers / eval features / argmax, gradients within tolerance).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(os.environ.get("RDP_VARIANT_CHILD") == "1", reason="already inside the variant run")
def test_variant_kernels_pass_the_parity_suite():
    env = dict(os.environ, RDP_PFN_ROWS="1", RDP_BWD_STREAM="1", RDP_VARIANT_CHILD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
