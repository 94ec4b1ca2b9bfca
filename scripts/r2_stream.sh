# usage (on a 4-GPU box): bash scripts/r2_stream.sh <tag>  - BASELINE config 5 as stated (one 65 536-image stream, LPT over the ranks) at N = 4 and 2
T=${1:-r2st}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/${T}_box.txt; nproc >> gpurun_out/${T}_box.txt
timeout 500 $TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 --workload config5 --stream 65536 > gpurun_out/${T}_stream_n4.json 2> gpurun_out/${T}_stream_n4.err
timeout 500 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --workload config5 --stream 65536 > gpurun_out/${T}_stream_n2.json 2> gpurun_out/${T}_stream_n2.err
echo done
