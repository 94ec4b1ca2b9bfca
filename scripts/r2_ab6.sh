# usage (on the GPU box): bash scripts/r2_ab6.sh <tag> - write pass with the look-back deferred to the end of a round and the
# slot-array flush: GPU tests, then the device-resident bench per stage against the previous structure (build/ab/lib_c_w6.so,
# lib_w4f1.so) and with 6 / 8 symbols per round (lib_n6.so, lib_n8.so); ncu source counters of the write kernel
tag=${1:-ab6}
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
  for lib in $IT build/ab/lib_n6.so build/ab/lib_n8.so build/ab/lib_c_w6.so; do run $lib --workload config2; done
done
run $IT --workload config3 --batch 1
run $IT --workload config4 --batch 1
run build/ab/lib_n8.so --workload config3 --batch 1
run $IT --workload config5
cat $out
[ -n "$SKIP_NCU" ] && exit 0
ncu --set full --import-source on --clock-control none -k regex:"k_huff_write" -c 1 -o gpurun_out/prof_${tag} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-cli --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv -k regex:k_huff_write --launch-skip 0 --launch-count 1 > gpurun_out/src_${tag}_k_huff_write.csv 2>/dev/null || true
rm -f gpurun_out/prof_${tag}.ncu-rep
echo done
