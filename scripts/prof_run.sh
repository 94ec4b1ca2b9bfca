# usage (on the GPU box): bash scripts/prof_run.sh <tag>   -> gpurun_out/<tag>_*
set -e
T=${1:-run}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.log || true
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_unstuff|k_huff_sync|k_huff_write|k_idct_color" -c 7 -o gpurun_out/prof_${T} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/prof_${T}.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:k_huff_sync --launch-skip 0 --launch-count 1 > gpurun_out/src_${T}_sync.csv 2>/dev/null || true
ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:k_huff_write > gpurun_out/src_${T}_write.csv 2>/dev/null || true
ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:k_idct_color > gpurun_out/src_${T}_idct.csv 2>/dev/null || true
