set -e
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s6_pytest.log
python bench.py --steps 2 --warmup 1 --batch 1024 --no-cpu-baseline --no-e2e > gpurun_out/s6_small.json 2> gpurun_out/s6.err
ncu --set full --import-source on --clock-control none -k regex:"k_huff_sync|k_huff_write|k_idct_color|k_unstuff" -c 10 -o gpurun_out/prof_s6 -f python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline --no-e2e > gpurun_out/s6_ncu.log 2>&1
ncu -i gpurun_out/prof_s6.ncu-rep --page raw --csv > gpurun_out/prof_s6_raw.csv
ncu -i gpurun_out/prof_s6.ncu-rep --page source --csv -k regex:k_huff_sync --launch-skip 0 --launch-count 1 > gpurun_out/src_s6_sync.csv 2>/dev/null || true
ncu -i gpurun_out/prof_s6.ncu-rep --page source --csv -k regex:k_huff_write > gpurun_out/src_s6_write.csv 2>/dev/null || true
