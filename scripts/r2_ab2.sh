# usage (on the GPU box): bash scripts/r2_ab2.sh <tag> [other.so ...] - entropy parity tests with every build, then the device-resident
# bench on config 2, in-tree against the other builds, twice, interleaved
tag=${1:-ab}; shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_ab.txt; : > $out
for lib in "" "$@"; do
  echo "== tests lib=${lib:-in-tree}" >> $out
  B200JPEG_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -2 >> $out
done
for rep in 1 2; do
  for lib in "" "$@"; do
    for wl in "config2" "config4 --batch 1"; do
      echo "== rep $rep lib=${lib:-in-tree} $wl" >> $out
      B200JPEG_LIB=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
    done
  done
done
cat $out
