# usage (on the GPU box): bash scripts/r2_run4.sh <tag>   - 16 KB K0 tiles, out-of-line pre-roll, TMA IDCT kernel, hybrid host
T=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --steps 10 --warmup 3 --no-cpu-baseline --no-cli --no-e2e > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
for i in 1 2; do
run tma_$i "X=1" ""
run notma_$i "B200JPEG_IDCT_TMA=0" ""
run pre512_$i "X=1" "--sync-preroll 512"
run pre1024_$i "X=1" "--sync-preroll 1024"
run pre1536_$i "X=1" "--sync-preroll 1536"
done
run pre1024_c5 "X=1" "--workload config5 --sync-preroll 1024"
run pre0_c5 "X=1" "--workload config5"
run notma_c5 "B200JPEG_IDCT_TMA=0" "--workload config5"
run pre1024_c4 "X=1" "--workload config4 --sync-preroll 1024"
run pre0_c4 "X=1" "--workload config4"
run pre0_c3 "X=1" "--workload config3"
run c3b1 "X=1" "--workload config3 --batch 1"
run c4b1 "X=1" "--workload config4 --batch 1"
timeout 600 python bench.py --steps 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
echo done
