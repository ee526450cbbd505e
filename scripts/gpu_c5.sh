# Config 4, config 5 (brute-force scale sweep) and C3 on one GPU, plus the isolated hot-loop microbenchmark.
mkdir -p gpurun_out
./raytrace_clj_b200/csrc/loopbench > gpurun_out/loopbench.log 2>&1; cat gpurun_out/loopbench.log
run() { name=$1; shift
  timeout 1200 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -c 300 gpurun_out/$name.err
  python - gpurun_out/$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); k=d["roofline"]["dominant_kernel"]
    print(sys.argv[1], "ms %.2f"%d["ms_per_step"], "value %.1fM samples/s"%(d["value"]/1e6), "tests %.3fT/s"%(d["tests_per_sec"]/1e12), "frac %.4f"%d["roofline"]["frac"], "cull frac %.3f"%k["frac"], "rays/sample %.3f surv/ray %.2f"%(d["rays_per_sample"], d["cull_survivors_per_ray"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), flush=True)
except Exception as e: print("no json:", e)
PY
}
run bench_c4 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline
run bench_c5_100 --workload c5-100 --steps 3 --warmup 3 --no-cpu-baseline
run bench_c5_1k --workload c5-1k --steps 3 --warmup 3 --no-cpu-baseline
run bench_c5_10k --workload c5-10k --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1
run bench_c3_strong_n1 --workload c3 --scaling strong --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1
run bench_c5_100k --workload c5-100k --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1
