# usage (on an N-GPU box): bash scripts/r2_multi.sh <tag> <ngpus>
T=${1:-r2m}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_box.txt; nproc >> gpurun_out/${T}_box.txt; free -g >> gpurun_out/${T}_box.txt; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name" >> gpurun_out/${T}_box.txt; df -h /dev/shm >> gpurun_out/${T}_box.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_n${N}.json 2> gpurun_out/${T}_n${N}.err
timeout 400 $TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --workload config5 --stream 65536 > gpurun_out/${T}_stream_n${N}.json 2> gpurun_out/${T}_stream_n${N}.err
if [ "$N" -gt 2 ]; then
timeout 300 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cli > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err
fi
if [ "$N" -gt 4 ]; then
timeout 300 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --no-cli > gpurun_out/${T}_n4.json 2> gpurun_out/${T}_n4.err
fi
timeout 300 python scripts/d2h_probe_multi.py > gpurun_out/${T}_d2h_probe.txt 2>&1
# one process, all GPUs (bj_create_multi): the CLI on 8192 files, against the same on one GPU
python - <<'PY' > gpurun_out/${T}_cli_multi.txt 2>&1
import os, sys, subprocess, time, shutil
sys.path.insert(0, "tests")
import jpeg_synth as js
d = "/dev/shm/bjcli"; shutil.rmtree(d, ignore_errors=True); os.makedirs(d)
uniq = [js.synth_jpeg(500, 375, seed=70000 + k, subsampling=2) for k in range(256)]
files = []
for i in range(8192):
    p = f"{d}/i{i:05d}.jpg"; open(p, "wb").write(uniq[i % 256]); files.append(p)
exe = "pim_jpeg_decoder_b200/host/_build/decoder_b200"
for ndev in (0, 1):
    env = dict(os.environ); 
    if ndev: env["B200JPEG_DEVICES"] = str(ndev)
    t = time.perf_counter(); out = subprocess.run([exe] + files, env=env, capture_output=True, text=True); dt = time.perf_counter() - t
    print("devices", ndev or "all", "wall %.2f s" % dt, "->", 8192 * 187500 / dt / 1e6, "Mpx/s"); print(out.stdout[-900:]); print(out.stderr[-300:])
shutil.rmtree(d, ignore_errors=True)
PY
echo done
