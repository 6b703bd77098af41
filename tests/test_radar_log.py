"""Consumer side of the estimator's CSV log: the reader that
mimo_precoder::compute_radar_aided_steering applies to the file, against files written in the reference's
format and (when oracle/_ref is built) by the reference's own range_angle_estimator."""
import os
import subprocess

import numpy as np

import mimo_ofdm_jrc as jrc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reader_on_reference_format(tmp_path):
    p = tmp_path / "radar_log.csv"
    assert jrc.radar_log_read_last(str(p)) is None                     # missing file
    p.write_text("")
    assert jrc.radar_log_read_last(str(p)) is None                     # empty file
    p.write_text("\n NEW RECORD - 10-17-2026 09:30:00\n")
    assert jrc.radar_log_read_last(str(p)) is None                     # header only
    with open(p, "a") as f:                                            # lib/range_angle_estimator_impl.cc:264-271
        f.write("09:30:00.125, \t0.0123, \t31.5, \t7.51, \t-12.5\n")
        f.write("09:30:00.375, \t0.0456, \t33.25, \t7.66, \t14.4775\n")
    t, power, snr, rng, ang = jrc.radar_log_read_last(str(p))
    assert t == "09:30:00.375" and power == 0.0456 and snr == 33.25 and rng == 7.66 and ang == 14.4775
    a = jrc.radar_aided_steering_vector(ang, 4)
    ref = np.exp(1j * np.pi * np.sin(np.deg2rad(14.4775)) * np.arange(4))
    np.testing.assert_allclose(a, ref, atol=1e-6)
    assert a.dtype == np.complex64 and abs(a[0] - 1) == 0


def test_reader_on_file_written_by_the_reference_estimator():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_vs_oracle")
    if not os.path.exists(exe):
        import pytest
        pytest.skip("oracle/_ref not built")
    log = "/tmp/jrc_ref_log.csv"
    if os.path.exists(log):
        os.remove(log)
    subprocess.run([exe], check=True, stdout=subprocess.DEVNULL, timeout=600)
    rec = jrc.radar_log_read_last(log)                                  # written by the reference's own code
    assert rec is not None and len(rec) == 5
    assert -90.0 <= rec[4] <= 90.0 and rec[3] >= 0.0 and np.isfinite(rec[1]) and np.isfinite(rec[2])
    txt = open(log).read()
    assert "NEW RECORD" in txt and txt.endswith("\n")
