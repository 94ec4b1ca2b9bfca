# usage (on the GPU box): bash scripts/e2e_run.sh <tag>
set -e
T=${1:-e2e}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_default.json 2> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --host-threads 8 > gpurun_out/${T}_ht8.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sub-batch-mb 12 > gpurun_out/${T}_sb12.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sub-batch-mb 48 --host-threads 8 > gpurun_out/${T}_sb48_ht8.json 2>> gpurun_out/${T}.err
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/${T}_smi.txt
nproc >> gpurun_out/${T}_smi.txt
