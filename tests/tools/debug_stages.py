"""GPU debug helper: compares every staged kernel with the oracle and reports the first mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np
import mimo_ofdm_jrc as jrc
from oracle import orc

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    eq = np.array_equal(a, b)
    if eq:
        print(f"{name}: identical"); return
    bad = np.argwhere(a != b)
    print(f"{name}: {len(bad)} / {a.size} differ, first at {bad[0]}, gpu {a[tuple(bad[0])]} ref {b[tuple(bad[0])]}, max abs diff {np.abs(a-b).max():.3e}")

rng = np.random.default_rng(0)
ch = jrc.Chain()
for n in (2, 8, 64, 128, 512, 1024, 4096, 16384):
    x = (rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))).astype(np.complex64)
    for fwd, sh in ((1, 0), (1, 1), (0, 0), (0, 1)):
        cmp(f"fft n={n} fwd={fwd} shift={sh}", ch.fft_vcc(x, fwd, sh), orc.fft_vcc(x, fwd, sh))
x = (rng.standard_normal((8, 512)) + 1j * rng.standard_normal((8, 512))).astype(np.complex64)
cmp("transpose 8x512 -> 512x128", ch.transpose_pad(x, 8, 16), orc.matrix_transpose(x, 8, 16))
x = (rng.standard_normal((32, 1024)) + 1j * rng.standard_normal((32, 1024))).astype(np.complex64)
cmp("transpose 32x1024 -> 1024x64", ch.transpose_pad(x, 32, 2), orc.matrix_transpose(x, 32, 2))
cmp("mag2", ch.mag_squared(x), orc.mag_squared(x))
