# usage (on the GPU box): bash scripts/e2e_run.sh <tag>   - A/B of the one-call path's knobs on one box
T=${1:-e2e}
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_smi100_$i.json 2> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clock-sample-ms 0 > gpurun_out/${T}_nosmi_$i.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clock-sample-ms 1000 > gpurun_out/${T}_smi1000_$i.json 2>> gpurun_out/${T}.err
done
