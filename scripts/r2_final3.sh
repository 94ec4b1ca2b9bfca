# usage (on the GPU box): bash scripts/r2_final3.sh <tag>  - the remaining one-GPU lines with the final kernels: config 5 as ONE
# 65 536-image stream, configs 3 / 4 as batches of 16 and config 4 as two images, the Level-0 (bj_exec_mcus) line
T=${1:-r2g}
mkdir -p gpurun_out
timeout 200 python bench.py --workload config5 --stream 65536 --no-cpu-baseline > gpurun_out/${T}_stream_n1.json 2> gpurun_out/${T}_stream_n1.err; tail -c 300 gpurun_out/${T}_stream_n1.json
for w in "config3" "config4" "config4 --batch 2"; do
  n=$(echo $w | tr -d ' -')
  timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cli > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
done
timeout 100 python bench.py --workload compat --steps 10 > gpurun_out/${T}_compat.json 2> gpurun_out/${T}_compat.err
echo done
