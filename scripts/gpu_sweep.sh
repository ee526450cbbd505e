mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
while read -r lanes ctas shape; do
  RT_WAVE_LANES=$lanes RT_CULL_CTAS_PER_SM=$ctas RT_CULL_SHAPE=$shape timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
  python - $lanes $ctas $shape <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("lanes/ctas/shape",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "cull %.2f (%.3f)"%(k["ms_per_step"],k["frac"]), {a:round(b,2) for a,b in k["other_stages_ms"].items()}, "e2e %.1fM"%(d["e2e"]["value"]/1e6), "surv %.3f"%d["cull_survivors_per_ray"], flush=True)
PY
done <<'CFG'
2 4 128x5
2 5 128x5
1 5 128x5
CFG
