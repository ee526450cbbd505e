mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "page_locked or identical_paths or trace_primary_new or render_new_primitive" > gpurun_out/r3_t2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r3_t2.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-c3 > gpurun_out/r3_b1.json 2> gpurun_out/r3_b1.err; echo "bench exit $?"; tail -2 gpurun_out/r3_b1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r3_b1.json") if l.startswith("{")][-1])
print("c2 ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], "e2e", d["e2e"]["value"])
PY
