# usage (on the GPU box): bash scripts/r2_run3.sh <tag>   - A/B: write-pass hand-over, sync pre-roll, per-image tickets
T=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --steps 10 --warmup 3 --no-cpu-baseline --no-cli --no-e2e > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
for i in 1 2; do
run new_$i "X=1" ""
run branchy_$i "B200JPEG_LIB=$PWD/build/ab/lib_branchy.so" ""
run pre256_$i "X=1" "--sync-preroll 256"
run pre512_$i "X=1" "--sync-preroll 512"
run pre1024_$i "X=1" "--sync-preroll 1024"
done
run pre512_c5 "X=1" "--workload config5 --sync-preroll 512"
run pre0_c5 "X=1" "--workload config5"
run pre512_c4 "X=1" "--workload config4 --sync-preroll 512"
run pre0_c4 "X=1" "--workload config4"
run pre512_c4b1 "X=1" "--workload config4 --batch 1 --sync-preroll 512"
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo done
