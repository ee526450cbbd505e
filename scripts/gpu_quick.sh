mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
while read -r kind entries; do
  RT_TAIL_KIND=$kind RT_TAIL_ENTRIES=$entries timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
  python - $kind $entries <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("tail kind/entries",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), "launches", d["gpu_launches"], "rays/sample %.5f"%d["rays_per_sample"], flush=True)
PY
done <<'CFG'
0 65536
1 65536
1 32768
1 131072
1 262144
CFG
