mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
while read -r claims lb wl steps; do
  RT_CULL_CLAIMS=$claims RT_LIGHT_BLOCK=$lb timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err || tail -3 gpurun_out/bench_s.err
  python - $claims $lb $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("claims/lightblock/workload",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), "launches", d["gpu_launches"], "rays/sample %.5f"%d["rays_per_sample"], flush=True)
PY
done <<'CFG'
0 256 c2 20
1 256 c2 20
1 128 c2 20
2 128 c2 20
4 128 c2 20
0 256 c3-slice 5
1 256 c3-slice 5
1 128 c3-slice 5
2 128 c3-slice 5
4 128 c3-slice 5
CFG
