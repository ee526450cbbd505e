mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_q.json")); print(d["ms_per_step"], d["roofline"]["frac"], d["cull_survivors_per_ray"], d["roofline"]["dominant_kernel"], d["e2e"]["value"])
PY
