# usage (on the GPU box): bash scripts/r2_run1.sh <tag>   - round 2, first pass: tests, bench, single-image latency, launch lists
T=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_gpus.txt; nproc >> gpurun_out/${T}_gpus.txt; free -g >> gpurun_out/${T}_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
B200JPEG_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --e2e-steps 2 > gpurun_out/${T}_trace.json 2> gpurun_out/${T}_trace.txt
timeout 300 python bench.py --no-cpu-baseline --steps 5 --staged-inputs > gpurun_out/${T}_staged.json 2> gpurun_out/${T}_staged.err
for w in "config3 --batch 1" "config3" "config4 --batch 1" "config4" "config5"; do
  n=$(echo $w | tr -d ' -')
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-cli > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_c3b1.csv python bench.py --workload config3 --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_c3b1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
timeout 200 python bench.py --workload compat --steps 10 > gpurun_out/${T}_compat.json 2> gpurun_out/${T}_compat.err
echo done
