mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
while read -r wl steps; do
  timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err || tail -3 gpurun_out/bench_s.err
  python - $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("workload",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "cull %.2f (%.3f)"%(k["ms_per_step"],k["frac"]), {a:round(b,2) for a,b in k["other_stages_ms"].items()}, "e2e %.1fM"%(d["e2e"]["value"]/1e6), flush=True)
PY
done <<'CFG'
c2 30
c4 5
CFG
