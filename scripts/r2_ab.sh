# usage (on the GPU box): bash scripts/r2_ab.sh <tag> [other.so ...] - GPU tests with the in-tree build, then the device-resident
# bench (config 3 / 4 with one image, config 2), in-tree against the other builds, twice, interleaved
tag=${1:-ab}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.txt 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.txt
tail -3 gpurun_out/${tag}_tests.txt
out=gpurun_out/${tag}_ab.txt; : > $out
for rep in 1 2; do
  for lib in "" "$@"; do
    for wl in "config3 --batch 1" "config4 --batch 1" "config2" "config5"; do
      echo "== rep $rep lib=${lib:-in-tree} $wl" >> $out
      B200JPEG_LIB=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
    done
  done
done
cat $out
