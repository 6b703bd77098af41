"""GPU: the stand-alone tcgen05 checks under scripts/ubench (hand-written UMMA descriptors, tf32 tile with the A operand
in shared memory and in tensor memory) compile for sm_100a and match a float64 reference."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UB = os.path.join(ROOT, "scripts", "ubench")


@pytest.mark.gpu
@pytest.mark.parametrize("name,marker", [("umma_tf32", "UMMA TF32 TILE OK"), ("umma_tmem_a", "UMMA TF32 TILE (A IN TMEM) OK")])
def test_tcgen05_tile(tmp_path, name, marker):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / name)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe, os.path.join(UB, name + ".cu")],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and marker in r.stdout, r.stdout + r.stderr
