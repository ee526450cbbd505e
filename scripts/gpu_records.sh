# Record runs for DESIGN.md's result table (1 GPU): every BASELINE config, brute force and BVH where it matters.
mkdir -p gpurun_out
run() { python bench.py --no-strong-c3 "$@" 2>/dev/null | grep '^{' | tail -1; }
run --steps 20 --warmup 5 --workload c1 --no-cpu-baseline > gpurun_out/r02_bench_c1.json
run --steps 10 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/r02_bench_c4.json
run --steps 5 --warmup 3 --workload c5-100 --no-cpu-baseline > gpurun_out/r02_bench_c5_100.json
run --steps 3 --warmup 3 --workload c5-1k --no-cpu-baseline > gpurun_out/r02_bench_c5_1k.json
run --steps 2 --warmup 3 --workload c5-10k --no-cpu-baseline > gpurun_out/r02_bench_c5_10k.json
run --steps 1 --warmup 3 --workload c5-100k --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_bench_c5_100k.json
run --steps 5 --warmup 3 --workload c5-100k --no-cpu-baseline --accel bvh > gpurun_out/r02_bench_c5_100k_bvh.json
run --steps 10 --warmup 3 --workload c2 --no-cpu-baseline --accel bvh > gpurun_out/r02_bench_c2_bvh.json
run --steps 10 --warmup 3 --workload c2 --no-cpu-baseline --variant 0 > gpurun_out/r02_bench_c2_megakernel.json
run --steps 5 --warmup 3 --workload cornell --no-cpu-baseline > gpurun_out/r02_bench_cornell.json
run --steps 5 --warmup 3 --workload final --no-cpu-baseline > gpurun_out/r02_bench_final.json
for f in gpurun_out/r02_bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f'.split('/')[-1], 'ms %.3f'%d['ms_per_step'], 'samples/s %.4g'%d['value'], 'tests/s %.4g'%d['tests_per_sec'], 'frac %.4f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'])"; done
