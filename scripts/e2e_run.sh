# usage (on the GPU box): bash scripts/e2e_run.sh <tag>
T=${1:-e2e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.log
for i in 1 2; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_direct_$i.json 2> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --staged-inputs > gpurun_out/${T}_staged_$i.json 2>> gpurun_out/${T}.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --host-threads 2 > gpurun_out/${T}_direct_ht2.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --host-threads 1 > gpurun_out/${T}_direct_ht1.json 2>> gpurun_out/${T}.err
