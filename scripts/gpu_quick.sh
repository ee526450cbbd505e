mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
for cfg in "2 4" "1 5"; do
  set -- $cfg
  RT_WAVE_LANES=$1 RT_CULL_CTAS_PER_SM=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q_$1_$2.json 2> gpurun_out/bench_q.err || tail -5 gpurun_out/bench_q.err
  python - $1 $2 <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_q_{sys.argv[1]}_{sys.argv[2]}.json")); k=d["roofline"]["dominant_kernel"]
print("lanes/ctas",sys.argv[1],sys.argv[2], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "cull %.2f (%.3f)"%(k["ms_per_step"],k["frac"]), k["other_stages_ms"], "e2e %.1fM"%(d["e2e"]["value"]/1e6), "launches", d["gpu_launches"], "surv/ray %.3f rays/sample %.4f"%(d["cull_survivors_per_ray"], d["rays_per_sample"]))
PY
done
