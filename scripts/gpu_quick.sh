mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
while read -r tail_entries wl; do
  RT_TAIL_ENTRIES=$tail_entries timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err || tail -3 gpurun_out/bench_s.err
  python - $tail_entries $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("tail entries/workload",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), "launches", d["gpu_launches"], flush=True)
PY
done <<'CFG'
65536 c2
131072 c2
262144 c2
524288 c2
1048576 c2
65536 c4
262144 c4
CFG
