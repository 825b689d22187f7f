import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    # GR_EMU_AS_CUDA=1: the `-m gpu` tests drive the CPU-emulated build of the CUDA library
    # (tests/emu/_build/libgenrich_emu.so, see tests/test_emu_library.py) instead of the real one.
    # A development aid for boxes without a GPU; the driver's GPU run never sets it.
    if os.environ.get("GR_EMU_AS_CUDA"):
        import subprocess
        from genrich_b200 import capi
        emu = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", emu, "_build/libgenrich_emu.so"])
        # GR_EMU_LIB: another build of it, e.g. one compiled with -fsanitize=address (run python under
        # LD_PRELOAD=libasan.so then): device buffers get red zones, kernels' out-of-bounds accesses abort
        capi._cuda_api = capi.Api(os.environ.get("GR_EMU_LIB") or os.path.join(emu, "_build", "libgenrich_emu.so"), "gr_")
