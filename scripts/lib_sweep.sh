# usage (on the GPU box): bash scripts/lib_sweep.sh <tag> <lib.so> [...]   - the device-resident bench with each build, twice, interleaved
T=${1:-sweep}; shift
mkdir -p gpurun_out
for i in 1 2; do
  for L in "$@"; do
    n=$(basename $L .so)
    B200JPEG_LIB=$PWD/$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_${n}_$i.json 2>> gpurun_out/${T}.err
  done
done
