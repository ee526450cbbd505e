mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
run() { name=$1; shift
  timeout 1200 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"
  python - gpurun_out/$name.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); k=d["roofline"]["dominant_kernel"]
print(sys.argv[1], "ms %.2f"%d["ms_per_step"], "value %.1fM samples/s"%(d["value"]/1e6), "tests %.3fT/s"%(d["tests_per_sec"]/1e12), "frac %.4f"%d["roofline"]["frac"], "cull frac %.3f"%k["frac"], flush=True)
PY
}
run bench_c5_10k --workload c5-10k --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1
run bench_c5_100k --workload c5-100k --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1
