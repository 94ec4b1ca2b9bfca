# usage (on the GPU box): bash scripts/e2e_run.sh <tag>
T=${1:-e2e}
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_default_$i.json 2> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ramp > gpurun_out/${T}_noramp_$i.json 2>> gpurun_out/${T}.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --host-threads 2 > gpurun_out/${T}_ht2.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sub-batch-mb 32 > gpurun_out/${T}_sb32.json 2>> gpurun_out/${T}.err
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/${T}_smi.txt
nproc >> gpurun_out/${T}_smi.txt; lscpu | grep -i "model name" >> gpurun_out/${T}_smi.txt
