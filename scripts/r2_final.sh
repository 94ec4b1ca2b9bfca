# usage (on the GPU box): bash scripts/r2_final.sh <tag>  - everything the round's profiles/ are made from (one GPU)
T=${1:-r2final}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_box.txt; nproc >> gpurun_out/${T}_box.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference.json 2> gpurun_out/${T}_reference.err
B200JPEG_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --e2e-steps 2 > gpurun_out/${T}_trace.json 2> gpurun_out/${T}_trace.txt
for w in "config3 --batch 1" "config3" "config4 --batch 1" "config4 --batch 2" "config4" "config5"; do
  n=$(echo $w | tr -d ' -')
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --no-cli > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
done
timeout 200 python bench.py --workload compat --steps 10 > gpurun_out/${T}_compat.json 2> gpurun_out/${T}_compat.err
timeout 900 python bench.py --workload config5 --stream 65536 --no-cpu-baseline > gpurun_out/${T}_stream_n1.json 2> gpurun_out/${T}_stream_n1.err
B200JPEG_IDCT_TMA=1 timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/${T}_idct_tma.json 2> gpurun_out/${T}_idct_tma.err
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e --streams 2 > gpurun_out/${T}_streams2.json 2> gpurun_out/${T}_streams2.err
timeout 300 python bench.py --steps 5 --no-cpu-baseline --staged-inputs > gpurun_out/${T}_staged.json 2> gpurun_out/${T}_staged.err
python scripts/d2h_probe_multi.py > gpurun_out/${T}_d2h_probe.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_c3b1.csv python bench.py --workload config3 --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_c3b1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_unstuff|k_subseq|k_huff_sync|k_huff_write|k_zero_tail|k_dc_predict|k_idct_color" -c 8 -o gpurun_out/prof_${T} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/prof_${T}.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
for k in k_unstuff k_huff_sync k_huff_write k_idct_color; do
  ncu -i gpurun_out/prof_${T}.ncu-rep --page source --csv -k regex:$k --launch-skip 0 --launch-count 1 > gpurun_out/src_${T}_$k.csv 2>/dev/null || true
done
rm -f gpurun_out/prof_${T}.ncu-rep
echo done
