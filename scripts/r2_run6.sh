# usage (on the GPU box): bash scripts/r2_run6.sh <tag>   - K0 with swizzled staging at 4 / 5 / 6 CTAs per SM
T=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --steps 10 --warmup 3 --no-cpu-baseline --no-cli --no-e2e > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
for i in 1 2; do
run cta6_$i "X=1" ""
run cta5_$i "B200JPEG_LIB=$PWD/build/ab/lib_unstuff5.so" ""
run cta4_$i "B200JPEG_LIB=$PWD/build/ab/lib_unstuff4.so" ""
done
run cta6_c5 "X=1" "--workload config5"
run cta6_c3b1 "X=1" "--workload config3 --batch 1"
run cta6_c4b1 "X=1" "--workload config4 --batch 1"
timeout 600 python bench.py --steps 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo done
