# A/B of the JRC_FUSED=1 wiring (tests/cpp/latency_blocks.cc): pieces of the transposed array's copy out, order of the radar
# block's own output kernel and the continuation's graph launch, short chain; prints p50 / p99 and the three blocks' p50 (us)
L=./gr-mimo-ofdm-jrc_b200/build/latency_blocks
ext() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.items():
    if 'JRC_FUSED' in k: print('$1', k[18:26], 'locked' if 'page-locked' in k else 'pageable', v['p50_us'], v['p99_us'], v['p50_us_mimo_ofdm_radar'], v['p50_us_matrix_transpose'], v['p50_us_range_angle_estimator'])
"; }
for rep in 1 2; do
for sh in ${SHORT:-0 1}; do for c in ${CHUNKS:-1 2 4}; do for pf in ${PADFIRST:-0 1}; do
JRC_FUSED_SHORT=$sh JRC_FUSED_CHUNKS=$c JRC_FUSED_PAD_FIRST=$pf timeout 60 $L 1200 2>/dev/null | ext "short=$sh chunks=$c padfirst=$pf"
done; done; done; done
