#!/bin/bash
# Builds oracle/_ref from the REFERENCE'S OWN SOURCES, compiled where they lie under /root/reference
# (nothing is copied into this repository).  The reference needs GNU Radio 3.8, Boost, Eigen, VOLK and
# FFTW, none of which exist in the build image, so the sources are compiled against header-only
# stand-ins: gr-mimo-ofdm-jrc_b200/gr_shim (runtime: block, tagged_stream_block, tags, pmt) and
# oracle/ref_shim (circular_buffer, the Eigen CSV formatter, three VOLK calls, gr::fft::fft_complex).
# Outputs: oracle/_ref/ref_vs_oracle (pins the oracle, run by tests/test_oracle_vs_ref.py) and
# oracle/_ref/libjrc_ref.so (reference chain for bench.py's CPU baseline).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${JRC_REFERENCE_DIR:-/root/reference}"
[ -d "$REF/lib" ] || { echo "reference tree not found at $REF"; exit 0; }
OUT="$HERE/_ref"
mkdir -p "$OUT"
make -C "$HERE" libjrc_oracle.so >/dev/null
CXX="${CXX:-g++}"
SRCS=(mimo_ofdm_radar_impl.cc matrix_transpose_impl.cc range_angle_estimator_impl.cc fft_peak_detect_impl.cc
      zero_pad_impl.cc target_simulator_impl.cc ofdm_cyclic_prefix_remover_impl.cc utils.cc)
FLAGS=(-O2 -std=c++14 -fPIC -w -ffp-contract=off
       -include chrono -include iomanip -include sstream -include fstream -include ctime -include cmath -include iostream
       -I"$HERE/ref_shim" -I"$HERE/../gr-mimo-ofdm-jrc_b200/gr_shim" -I"$REF/include" -I"$REF/lib" -I"$HERE")
OBJS=()
for s in "${SRCS[@]}"; do
    o="$OUT/${s%.cc}.o"
    "$CXX" "${FLAGS[@]}" -c "$REF/lib/$s" -o "$o"
    OBJS+=("$o")
done
"$CXX" "${FLAGS[@]}" -shared -o "$OUT/libjrc_ref.so" "$HERE/ref_harness.cc" "${OBJS[@]}" -L"$HERE" -ljrc_oracle -Wl,-rpath,'$ORIGIN/..'
"$CXX" "${FLAGS[@]}" -DREF_HARNESS_MAIN -o "$OUT/ref_vs_oracle" "$HERE/ref_harness.cc" "${OBJS[@]}" -L"$HERE" -ljrc_oracle -Wl,-rpath,'$ORIGIN/..'
rm -f "${OBJS[@]}"
echo "built $OUT/libjrc_ref.so and $OUT/ref_vs_oracle"
