# usage (on the GPU box): bash scripts/r2_ab7.sh <tag> - bit reader with a 32-bit word index, shorter magnitude extension: GPU
# tests, then the device-resident bench per stage against the build before (build/ab/lib_n6.so)
tag=${1:-ab7}
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
  for lib in $IT build/ab/lib_n6.so; do run $lib --workload config2; done
done
for lib in $IT build/ab/lib_n6.so; do run $lib --workload config3 --batch 1; run $lib --workload config4 --batch 1; run $lib --workload config5; done
cat $out
