"""The C++ GNU Radio block wrappers (gr-mimo-ofdm-jrc_b200/lib) driven through the runtime stand-in.

GPU: build/test_blocks feeds tagged packets to every block's general_work() and compares outputs,
tags, consumed counts and published messages with the oracle (bit-exact per block; the fused
radar_chain block within 1e-4 of the map peak).
CPU: the wrapper library links, exports the reference's make() symbols, and refuses to construct a
block without a CUDA device."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200")
EXE = os.path.join(PKG, "build", "test_blocks")
LIB = os.path.join(PKG, "libgnuradio-mimo_ofdm_jrc.so")


def _built():
    return os.path.exists(EXE) and os.path.exists(LIB)


def test_wrapper_library_exports_the_reference_factories():
    assert _built(), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    out = subprocess.run(["nm", "-DC", LIB], capture_output=True, text=True, check=True).stdout
    for blk in ("mimo_ofdm_radar", "matrix_transpose", "range_angle_estimator", "fft_peak_detect", "zero_pad", "radar_chain",
                "ofdm_cyclic_prefix_remover", "target_simulator"):
        assert f"gr::mimo_ofdm_jrc::{blk}::make(" in out, blk
    ctypes.CDLL(os.path.join(PKG, "libjrc_cuda.so"))
    ctypes.CDLL(LIB)


def test_blocks_refuse_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_cpp_blocks_against_oracle():
    assert _built()
    env = dict(os.environ, JRC_GOLDEN_DIR=os.path.join(ROOT, "tests", "golden"))     # + the capture_radar_data() format test
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "ALL BLOCK TESTS PASSED" in r.stdout


@pytest.mark.gpu
def test_streaming_latency_harness_runs():
    """configs[3]: the fused block driven one CPI per general_work() call (tests/cpp/latency_blocks.cc)."""
    import json
    exe = os.path.join(PKG, "build", "latency_blocks")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])      # (the blocks log to stdout too)
    assert len(res) == 8          # per map size: the five separate blocks, the same wiring with JRC_FUSED=1, the fused block, the pipelined fused block
    for k, v in res.items():
        if "pipeline" in k:
            assert 0 < v["latency_p50_us"] <= v["latency_p99_us"] < 20000 and v["sustained_cpi_per_s"] > 1000 and v["cpis"] == 300, k
            continue
        assert 0 < v["p50_us"] <= v["p99_us"] < 20000, k
        if k.startswith("radar_chain"):
            assert v["calls"] == 300
