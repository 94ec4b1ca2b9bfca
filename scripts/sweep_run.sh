T=${1:-sw}
mkdir -p gpurun_out
for s in 1 2 4 8; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --slices $s > gpurun_out/${T}_slices$s.json 2> gpurun_out/${T}.err
done
for b in 1536 2048 3072 4096 6144; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --subseq-bits $b > gpurun_out/${T}_bits$b.json 2>> gpurun_out/${T}.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --subseq-bits 4096 --slices 8 > gpurun_out/${T}_bits4096_s8.json 2>> gpurun_out/${T}.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --subseq-bits 2048 --slices 2 > gpurun_out/${T}_bits2048_s2.json 2>> gpurun_out/${T}.err
