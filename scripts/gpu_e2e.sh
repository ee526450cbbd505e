mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_c_abi.py -m gpu -x -q -k "page_locked or c_program or tail_warp" > gpurun_out/r3_t2.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r3_t2.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-strong-c3 > gpurun_out/r3_b1.json 2> gpurun_out/r3_b1.err; echo "bench exit $?"; tail -2 gpurun_out/r3_b1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r3_b1.json") if l.startswith("{")][-1])
print("c2 ms", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
PY
