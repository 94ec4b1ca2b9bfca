# usage (on the GPU box): bash scripts/ab_run.sh <tag> <other.so> [other2.so ...]
# A/B on one box: GPU tests with the in-tree library, then the device-resident bench alternating between the in-tree
# library and the given builds (B200JPEG_LIB), then per-kernel SASS profiles of the in-tree build.
T=${1:-ab}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
for i in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_new_$i.json 2> gpurun_out/${T}.err
  n=0
  for L in "$@"; do
    n=$((n+1))
    B200JPEG_LIB=$PWD/$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_other${n}_$i.json 2>> gpurun_out/${T}.err
  done
done
[ -n "$SKIP_NCU" ] && exit 0
ncu --set full --import-source on --clock-control none -k regex:"k_huff_sync|k_huff_write" -c 4 -o gpurun_out/prof_${T} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/prof_${T}.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:k_huff_sync --launch-skip 0 --launch-count 1 > gpurun_out/src_${T}_sync.csv 2>/dev/null || true
ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:k_huff_write > gpurun_out/src_${T}_write.csv 2>/dev/null || true
