# usage (on the GPU box): bash scripts/r2_final2.sh <tag>  - end-of-round evidence after the IDCT / write-pass / bit-reader changes (one GPU):
# GPU tests, smoke, one `ncu --set full` capture of a step (-> profiles/ncu_traffic.json for the kernel sources as they are), launch lists,
# the default bench (with e2e, reference on the host cores, parity, CLI), configs 3 / 4 as single images, config 5, a parity fuzz
T=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_box.txt; nproc >> gpurun_out/${T}_box.txt
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_unstuff|k_subseq|k_huff_sync|k_huff_write|k_zero_tail|k_dc_predict|k_idct_color" -c 8 -o gpurun_out/prof_${T} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-cli --streams 1 > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/prof_${T}.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
for k in k_unstuff k_huff_sync k_huff_write k_idct_color; do
  ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:$k --launch-skip 0 --launch-count 1 > gpurun_out/src_${T}_$k.csv 2>/dev/null || true
done
rm -f gpurun_out/prof_${T}.ncu-rep
python scripts/make_traffic_json.py gpurun_out/prof_${T}_raw.csv "profiles/r2_ncu_full_raw.csv (scripts/r2_final2.sh, one step of the default bench)" > /dev/null && cp profiles/ncu_traffic.json gpurun_out/${T}_ncu_traffic.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-cli > gpurun_out/${T}_ncu1.log 2>&1
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
for w in "config3 --batch 1" "config4 --batch 1" "config5"; do
  n=$(echo $w | tr -d ' -')
  timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cli > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
done
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${T}_launches_c3b1.csv python bench.py --workload config3 --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-cli > gpurun_out/${T}_ncu_c3b1.log 2>&1
timeout 100 python scripts/fuzz_gpu.py 3072 77 > gpurun_out/${T}_fuzz.txt 2>&1; tail -1 gpurun_out/${T}_fuzz.txt
echo done
