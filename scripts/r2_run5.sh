# usage (on the GPU box): bash scripts/r2_run5.sh <tag>   - multi-stream device-resident value, CLI I/O threads, stream 65536 on one GPU
T=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --steps 10 --warmup 3 --no-cpu-baseline --no-cli --no-e2e > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
run s1 "X=1" "--streams 1"
run s2 "X=1" "--streams 2"
run s3 "X=1" "--streams 3"
run s4 "X=1" "--streams 4"
run s2_c5 "X=1" "--workload config5 --streams 2"
run s3_c5 "X=1" "--workload config5 --streams 3"
run s2_c4 "X=1" "--workload config4 --streams 2"
run s2_c3 "X=1" "--workload config3 --streams 2"
timeout 600 python bench.py --steps 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python bench.py --workload config5 --stream 65536 --no-cpu-baseline > gpurun_out/${T}_stream_n1.json 2> gpurun_out/${T}_stream_n1.err
python scripts/d2h_probe_multi.py > gpurun_out/${T}_d2h_probe.txt 2>&1
echo done
