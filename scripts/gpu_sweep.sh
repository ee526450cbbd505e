mkdir -p gpurun_out
while read -r common wl steps; do
  RT_COMMON_ORIGIN=$common timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
  python - $common $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("common/workload",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "cull %.2f (%.3f)"%(k["ms_per_step"],k["frac"]), {a:round(b,2) for a,b in k["other_stages_ms"].items()}, flush=True)
PY
done <<'CFG'
1 c5-10k 1
0 c5-10k 1
CFG
