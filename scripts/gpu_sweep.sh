mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "render or counters or sharding" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err || tail -3 gpurun_out/bench_s.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_s.json")); print("c2 ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), flush=True)
PY
done
