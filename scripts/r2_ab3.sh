# usage (on the GPU box): bash scripts/r2_ab3.sh <tag> <lib.so> ... - device-resident bench on config 2 with each build, twice,
# interleaved; then the parity tests with the builds named in $TEST_LIBS
tag=${1:-ab}; shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_ab.txt; : > $out
for rep in 1 2; do
  for lib in "$@"; do
    echo "== rep $rep lib=$lib config2" >> $out
    B200JPEG_LIB=$lib timeout 300 python bench.py --workload config2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
  done
done
for lib in $TEST_LIBS; do
  echo "== tests lib=$lib" >> $out
  B200JPEG_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -2 >> $out
done
cat $out
