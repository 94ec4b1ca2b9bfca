# usage (on a multi-GPU box): bash scripts/multi_run.sh <tag> <ngpus>
T=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_n${N}.json 2> gpurun_out/${T}_n${N}.err
nproc > gpurun_out/${T}_nproc.txt; nvidia-smi topo -m >> gpurun_out/${T}_nproc.txt 2>&1; lscpu | grep -i -E "numa|socket|^CPU\(s\)" >> gpurun_out/${T}_nproc.txt
