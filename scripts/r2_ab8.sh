# usage (on the GPU box): bash scripts/r2_ab8.sh <tag> - write pass with the symbols of a round unrolled: GPU tests, device-resident
# bench per stage for 6 (in-tree), 5 and 8 symbols per round against the rolled loop (build/ab/lib_rolled6.so)
tag=${1:-ab8}
mkdir -p gpurun_out
out=gpurun_out/${tag}.txt; : > $out
echo "== tests in-tree" >> $out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> $out
run() { lib=$1; shift
  echo "== lib=$lib $*" >> $out
  B200JPEG_LIB=$lib timeout 300 python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
}
IT=pim_jpeg_decoder_b200/libb200jpeg.so
for rep in 1 2; do
  for lib in $IT build/ab/lib_u5.so build/ab/lib_u8.so build/ab/lib_rolled6.so; do run $lib --workload config2; done
done
run $IT --workload config3 --batch 1
run $IT --workload config5
cat $out
