"""Kernel variants behind environment knobs that were written where no GPU was at hand (checked on the
CPU by tests/emu, see tests/test_emu_kernels.py).  They are NOT the default path.  Each runs in a child
process with a time limit and is marked xfail(strict=False): until a variant has been seen on a device
its outcome is information (XPASS: same bits as the default path on the device; xfail: not yet), not a
gate -- the gates are the tests of the default path in the other files."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "rank512": ["GR_FUSED_RANK=1"],
    "rank1024": ["GR_FUSED_RANK=1", "GR_FR_CAP=1024"],
    "rank512_slots": ["GR_FUSED_RANK=1", "GR_FB_SLOTS=1"],
    "ue_warp": ["GR_UE_WARP=1"],
    "ur_groups4": ["GR_UR_GROUPS=4"],
    "cl_tiles4": ["GR_CL_TILES=4"],
    "p2": ["GR_FB_P2=1"],
    "all_p2": ["GR_FUSED_RANK=1", "GR_FB_P2=1", "GR_UE_WARP=1", "GR_UR_GROUPS=4", "GR_CL_TILES=4"],
    "all": ["GR_FUSED_RANK=1", "GR_FB_SLOTS=1", "GR_UE_WARP=1", "GR_UR_GROUPS=4", "GR_CL_TILES=4"],
}


@pytest.mark.xfail(reason="variant not yet run on a device (CPU-emulated only)", strict=False)
@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_same_bits_as_default(name):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variants_check.py")] + VARIANTS[name],
                       capture_output=True, text=True, timeout=420)
    print(p.stdout[-2000:], p.stderr[-4000:])
    assert p.returncode == 0
