# the round's closing run on one B200: GPU suite, smoke(), sanitizer on the block tests (fused-mode graphs), bench line, launch
# list of the latency harness; outputs under gpurun_out/
set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
export JRC_GOLDEN_DIR=$PWD/tests/golden
( echo "== memcheck build/test_blocks"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 ./gr-mimo-ofdm-jrc_b200/build/test_blocks 2>&1 | grep -v "^\[" | tail -4;
  echo "== racecheck build/test_blocks"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 ./gr-mimo-ofdm-jrc_b200/build/test_blocks 2>&1 | grep -v "^\[" | tail -4 ) > gpurun_out/r2_sanitizer_blocks.txt 2>&1
tail -12 gpurun_out/r2_sanitizer_blocks.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1_v9.json 2> gpurun_out/b9n1.err; tail -c 300 gpurun_out/r2_bench_n1_v9.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/lat_launches3.csv ./gr-mimo-ofdm-jrc_b200/build/latency_blocks 8 > /dev/null 2>&1
true
