"""The plain-C driver of the ABI (integration/c_driver.c): builds with gcc -std=c99 against include/deepfm_b200.h,
validates arguments without a GPU, trains / evaluates a small model from host buffers on one."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _exe():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "recommender_tensorflow_b200", "libdeepfm_b200.so")):
        g.build()
    return g.build_c_driver()


def test_c_driver_builds_and_validates_without_gpu():
    out = subprocess.run([_exe(), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C_DRIVER_OK" in out.stdout


@pytest.mark.gpu
def test_c_driver_trains_from_host_buffers():
    out = subprocess.run([_exe()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C_DRIVER_OK" in out.stdout
