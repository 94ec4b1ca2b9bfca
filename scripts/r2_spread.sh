# usage (on an 8-GPU box): bash scripts/r2_spread.sh <tag>  - does spreading 2 / 4 ranks over the 8 GPUs give them more copy-out bandwidth?
T=${1:-r2s}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4; do
  B200JPEG_BENCH_SPREAD=1 timeout 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --no-cli --e2e-steps 6 > gpurun_out/${T}_spread_n$n.json 2> gpurun_out/${T}_spread_n$n.err
  B200JPEG_BENCH_SPREAD=0 timeout 300 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 --no-cli --e2e-steps 6 > gpurun_out/${T}_packed_n$n.json 2> gpurun_out/${T}_packed_n$n.err
done
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
lspci -tv 2>/dev/null | head -80 >> gpurun_out/${T}_topo.txt
echo done
