# usage (on the GPU box): bash scripts/r2_run2.sh <tag>   - round 2, second pass: single-pass K0, REF_MCUS, latency experiments
T=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --steps 20 --warmup 5 --no-cpu-baseline --no-cli --no-e2e > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
run c3b1_default "X=1" "--workload config3 --batch 1"
run c3b1_whole "B200JPEG_RI_SPLIT=0" "--workload config3 --batch 1"
run c3b1_min64 "B200JPEG_MIN_SUB=64" "--workload config3 --batch 1"
run c3b1_min256 "B200JPEG_MIN_SUB=256" "--workload config3 --batch 1"
run c3b16_default "X=1" "--workload config3"
run c3b16_split "B200JPEG_RI_SPLIT=1000000" "--workload config3"
run c4b1_default "X=1" "--workload config4 --batch 1"
run c4b1_min64 "B200JPEG_MIN_SUB=64" "--workload config4 --batch 1"
run c4b1_min256 "B200JPEG_MIN_SUB=256" "--workload config4 --batch 1"
run c4b2_default "X=1" "--workload config4 --batch 2"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_c3b1.csv python bench.py --workload config3 --batch 1 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_c3b1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
echo done
