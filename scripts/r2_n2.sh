# usage (on a 2-GPU box): bash scripts/r2_n2.sh <tag>  - repeatability of the 2-rank end-to-end number
T=${1:-r2n2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; nproc >> gpurun_out/${T}_topo.txt
for i in 1 2 3; do
timeout 300 $TR --nproc-per-node 2 --master-port 2951$i bench.py --gpus 2 --steps 10 --warmup 3 --no-cli --e2e-steps 8 > gpurun_out/${T}_$i.json 2> gpurun_out/${T}_$i.err
done
echo done
