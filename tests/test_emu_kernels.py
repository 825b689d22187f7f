"""CUDA kernels on the CPU: tests/emu compiles kernels of genrich_b200/csrc *as they are* (extracted
by name from the .cu files) against a lock-step emulation of warps / CTAs (tests/emu/cuda_emu.h) and
compares kernels that have not been on a GPU yet with the ones validated there, and both with a plain
per-cell walk.  Test infrastructure only; the product path never runs on the CPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def _run(target, order=0):
    subprocess.check_call(["make", "-s", "-C", EMU, "_build/" + target])
    p = subprocess.run([os.path.join(EMU, "_build", target)], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, EMU_ORDER=str(order)))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


@pytest.mark.parametrize("order", [0, 2])
def test_fused_scan_rank_form_equals_validated_kernel(order):
    """k_fr_scan (GR_FUSED_RANK=1: warp-owned 8192-cell blocks, rank form, 64 / 512 / 1024 distinct
    cells per round) == k_fb_scan (the default, validated on the B200) == per-cell walk: interval
    ends, float bits, chromosome starts, break bitmap, error flags; edge inputs, hot spots,
    fractional weights, blocks that need several rounds, unsaved and foreign chromosomes.  Also the
    slot path (GR_FB_SLOTS=1: k_fb_move_slot + the gated exact chain), without and with overflow, and
    the two-level partition (GR_FB_P2=1: k_p1_count / scan / move + k_p2) == count -> scan -> move."""
    out = _run("emu_fused_scan", order)     # lanes resumed in order / in reverse / in a changing order
    assert "FAIL" not in out and out.count(" ok") == 6, out


@pytest.mark.parametrize("order", [0, 2])
def test_union_emit_warp_form_equals_validated_kernel(order):
    """k_union_emit_w (GR_UE_WARP=1: one warp per bitmap block) == k_union_emit<4> / <2> (validated on
    the B200) == a plain walk over the bits: interval ends, gathered pileup values, union bitmap,
    chromosome starts; empty, sparse and dense blocks (several list rounds), partial last CTA.
    Also k_union_rank_g<2|4> (GR_UR_GROUPS: 32 / 64 blocks per look-back tile) == k_union_rank == host ranks."""
    out = _run("emu_union", order)
    assert "FAIL" not in out and out.count(" ok") == 16, out


@pytest.mark.parametrize("order", [0, 2])
def test_ctrl_clamp_long_tiles_equals_validated_kernel(order):
    """k_ctrl_clamp_m<4> (GR_CL_TILES=4: 32768 raw intervals per look-back tile, 128-wide window) ==
    k_ctrl_clamp (validated on the B200) == a plain loop over savePileupCtrl's rule: clamped values,
    surviving boundaries, chromosome starts (chromosomes with one / no intervals / no slots), bitmap."""
    out = _run("emu_ctrl", order)
    assert "FAIL" not in out and out.count(" ok") == 6, out
