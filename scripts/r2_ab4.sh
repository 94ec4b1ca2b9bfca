# usage (on the GPU box): bash scripts/r2_ab4.sh <tag> - GPU tests with the in-tree library; device-resident bench per stage for the
# in-tree build against the builds under build/ab/; ncu source counters of the write and IDCT kernels of the in-tree build
tag=${1:-ab4}
mkdir -p gpurun_out
out=gpurun_out/${tag}.txt; : > $out
echo "== tests in-tree" >> $out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> $out
run() { # lib workload-args...
  lib=$1; shift
  echo "== lib=$lib $*" >> $out
  B200JPEG_LIB=$lib timeout 300 python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
}
IT=pim_jpeg_decoder_b200/libb200jpeg.so
for rep in 1 2; do
  for lib in $IT build/ab/lib_c1.so build/ab/lib_head.so; do run $lib --workload config2; done
done
run build/ab/lib_c_w6.so --workload config2
run build/ab/lib_c_w8.so --workload config2
for lib in $IT build/ab/lib_head.so; do run $lib --workload config4 --batch 1; run $lib --workload config3 --batch 1; done
cat $out
[ -n "$SKIP_NCU" ] && exit 0
ncu --set full --import-source on --clock-control none -k regex:"k_huff_write|k_idct_color" -c 2 -o gpurun_out/prof_${tag} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-cli --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
for k in k_huff_write k_idct_color; do
  ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv -k regex:$k --launch-skip 0 --launch-count 1 > gpurun_out/src_${tag}_$k.csv 2>/dev/null || true
done
rm -f gpurun_out/prof_${tag}.ncu-rep
echo done
