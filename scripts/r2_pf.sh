# usage (on the GPU box): bash scripts/r2_pf.sh - A/B of the bit reader's optional stream prefetch (-DBJ_STREAM_PREFETCH=<bytes>):
# one 4K image (config 3 / 4 as stated) and config 2, in-tree build against build/ab/lib_pf*.so, twice, interleaved
mkdir -p gpurun_out
out=gpurun_out/pf_ab.txt; : > $out
for rep in 1 2; do
  for lib in "" build/ab/lib_pf40.so build/ab/lib_pf72.so build/ab/lib_pf136.so; do
    for wl in "config3 --batch 1" "config4 --batch 1" "config2"; do
      echo "== rep $rep lib=${lib:-in-tree} $wl" >> $out
      B200JPEG_LIB=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
    done
  done
done
cat $out
