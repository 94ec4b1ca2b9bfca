T=${1:-c}
mkdir -p gpurun_out
for w in config3 config4 config5; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
done
python bench.py --workload config3 --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_config3_b1.json 2> gpurun_out/${T}_config3_b1.err
python bench.py --workload config4 --batch 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_config4_b2.json 2> gpurun_out/${T}_config4_b2.err
