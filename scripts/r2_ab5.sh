# (record of an experiment: applies to commit c8d30e7 "Synchronisation pass, tail mode", which was measured with this script and reverted - the B200JPEG_SYNC_TAIL switch does not exist at HEAD)
# usage (on the GPU box): bash scripts/r2_ab5.sh <tag> - tail mode of the synchronisation pass: GPU tests (default options, then the
# whole suite with tail mode forced), device-resident bench per stage with and without it, launch list of the sync kernels
tag=${1:-ab5}
mkdir -p gpurun_out
out=gpurun_out/${tag}.txt; : > $out
echo "== tests (default options)" >> $out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> $out
echo "== tests (B200JPEG_SYNC_TAIL=2: every decode in tail mode)" >> $out
B200JPEG_SYNC_TAIL=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -3 >> $out
run() { # env-value workload-args...
  v=$1; shift
  echo "== sync_tail=$v $*" >> $out
  B200JPEG_SYNC_TAIL=$v timeout 300 python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cli --streams 1 --clock-sample-ms 0 2>/dev/null | python scripts/bench_line.py >> $out
}
for rep in 1 2; do
  for v in 1 0; do run $v --workload config2; done
done
for v in 1 0; do run $v --workload config5; done
for v in 2 0; do run $v --workload config4 --batch 1; done
cat $out
[ -n "$SKIP_NCU" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_huff_sync|k_huff_write" -c 40 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-cli --streams 1 > gpurun_out/${tag}_ncu.log 2>&1
grep -E "k_huff" gpurun_out/${tag}_launches.csv | awk -F'","' '{print $5, $NF}' | tail -24
echo done
