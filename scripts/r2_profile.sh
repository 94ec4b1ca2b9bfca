# usage (on the GPU box): bash scripts/r2_profile.sh <tag>  - ncu evidence: launch list + one --set full capture of every kernel of a step
T=${1:-r2p}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_unstuff|k_subseq|k_huff_sync|k_huff_write|k_zero_tail|k_dc_predict|k_idct_color" -c 8 -o gpurun_out/prof_${T} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/prof_${T}.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
for k in k_unstuff k_huff_sync k_huff_write k_idct_color; do
  ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:$k --launch-skip 0 --launch-count 1 > gpurun_out/src_${T}_$k.csv 2>/dev/null || true
done
rm -f gpurun_out/prof_${T}.ncu-rep
echo done
