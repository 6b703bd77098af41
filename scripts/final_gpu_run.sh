set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export JRC_GOLDEN_DIR=$PWD/tests/golden
( echo "== memcheck build/test_blocks"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 ./gr-mimo-ofdm-jrc_b200/build/test_blocks 2>&1 | grep -v "^\[" | tail -4; echo "rc=$?";
  echo "== racecheck build/test_blocks"; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 ./gr-mimo-ofdm-jrc_b200/build/test_blocks 2>&1 | grep -v "^\[" | tail -4 ) > gpurun_out/r2_sanitizer_blocks.txt 2>&1
tail -12 gpurun_out/r2_sanitizer_blocks.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1_v8.json 2> gpurun_out/b8n1.err; tail -c 600 gpurun_out/r2_bench_n1_v8.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/lat_launches2.csv ./gr-mimo-ofdm-jrc_b200/build/latency_blocks 8 > /dev/null 2>&1
