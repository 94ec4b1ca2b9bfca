# usage (on the GPU box): bash scripts/final_run.sh <tag>  - everything the round's profiles/ are made from
T=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference.json 2> gpurun_out/${T}_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_unstuff|k_subseq|k_huff_sync|k_huff_write|k_zero_tail|k_dc_predict|k_idct_color" -c 8 -o gpurun_out/prof_${T} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/prof_${T}.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
for k in k_unstuff k_huff_sync k_huff_write k_idct_color; do
  ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:$k --launch-skip 0 --launch-count 1 > gpurun_out/src_${T}_$k.csv 2>/dev/null || true
done
for w in config3 config4 config5; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
done
