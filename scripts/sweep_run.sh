T=${1:-sw}
mkdir -p gpurun_out
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --subseq-bits $1 --slices $2 $3 > gpurun_out/${T}_b$1_s$2$4.json 2>> gpurun_out/${T}.err; }
run 0 0 "" _c2; run 0 0 "--workload config3" _c3; run 0 0 "--workload config3 --batch 1" _c3b1; run 0 0 "--workload config4" _c4; run 0 0 "--workload config4 --batch 2" _c4b2; run 0 0 "--workload config5" _c5
