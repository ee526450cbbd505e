# Round-2 records with the tensor-core cull (default) and the FP32-pipe A/B, C2 / C1 / C4 / C5-1k / cornell
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || { echo "$name FAILED"; tail -3 gpurun_out/$name.err; }; python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/$name.json") if l.startswith("{")][-1])
    print("$name", round(d["ms_per_step"],3), "ms  frac", round(d["roofline"]["frac"],4), " e2e ms", round(d["e2e"]["ms_per_step"],3), " dom", d["roofline"]["dominant_kernel"] and round(d["roofline"]["dominant_kernel"]["frac"],3), " tensor", d["roofline"].get("tensor") and round(d["roofline"]["tensor"]["frac"],3))
except Exception as e: print("$name", "parse error", e)
P
}
run r02_bench_c2_tc --steps 20 --warmup 5
run r02_bench_c2_fp32cull --steps 20 --warmup 5 --cull fp32 --no-strong-c3 --no-cpu-baseline
run r02_bench_c1_tc --workload c1 --steps 20 --warmup 5 --no-cpu-baseline
run r02_bench_c4_tc --workload c4 --steps 10 --warmup 3 --no-cpu-baseline
run r02_bench_c5_1k_tc --workload c5-1k --steps 5 --warmup 3 --no-cpu-baseline
run r02_bench_c5_100_tc --workload c5-100 --steps 10 --warmup 3 --no-cpu-baseline
run r02_bench_cornell_tc --workload cornell --steps 10 --warmup 3 --no-cpu-baseline
run r02_bench_c5_10k_tc --workload c5-10k --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c5_100k_tc --workload c5-100k --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1
