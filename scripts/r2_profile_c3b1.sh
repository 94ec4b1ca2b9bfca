# usage (on the GPU box): bash scripts/r2_profile_c3b1.sh  - ncu --set full of the Huffman kernels on ONE 4K image (config 3 as stated)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_unstuff|k_huff_sync|k_huff_write|k_idct_color" -c 6 -o gpurun_out/prof_c3b1 -f python bench.py --workload config3 --batch 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --streams 1 > gpurun_out/c3b1_ncu.log 2>&1
ncu -i gpurun_out/prof_c3b1.ncu-rep --page raw --csv > gpurun_out/prof_c3b1_raw.csv
ncu -i gpurun_out/prof_c3b1.ncu-rep --page source --csv -k regex:k_huff_sync --launch-skip 0 --launch-count 1 > gpurun_out/src_c3b1_sync.csv 2>/dev/null || true
rm -f gpurun_out/prof_c3b1.ncu-rep
echo done
