"""mimo_ofdm_jrc -- B200-native radar hot path behind the reference's block names.

Mirrors `import mimo_ofdm_jrc` of the reference (python/__init__.py:35-39,
swig/mimo_ofdm_jrc_swig.i:31-68) for the blocks on the radar path.  Every block calls
libjrc_cuda.so through ctypes (cabi.py); nothing is computed in Python.
"""
from .cabi import Chain, JrcError, DET_DTYPE, PATH_AUTO, PATH_FUSED, PATH_STAGED, PATH_TILED, DET_PASSED, DET_EXACT, LIB_PATH, EXPORTS  # noqa: F401
from .blocks import (mimo_ofdm_radar, matrix_transpose, range_angle_estimator, fft_peak_detect,  # noqa: F401
                     zero_pad, fft_vcc, complex_to_mag_squared, radar_chain, ofdm_cyclic_prefix_remover, nlog10_ff, target_simulator,
                     radar_log_read_last, radar_aided_steering_vector)
from . import synth, shard  # noqa: F401
