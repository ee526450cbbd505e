# final state at N GPUs: the bench as the driver launches it (weak C2 + in-library e2e + strong_c3 block) and the multi-device tests
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r02_final_bench_c2_n$N.json 2> gpurun_out/r02_final_bench_c2_n$N.err
echo "bench exit $?"; tail -c 300 gpurun_out/r02_final_bench_c2_n$N.err
python - gpurun_out/r02_final_bench_c2_n$N.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "ms %.3f"%d["ms_per_step"], "value %.1fM samples/s"%(d["value"]/1e6), "frac %.4f"%d["roofline"]["frac"], "e2e %.1fM (%.3f ms)"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), d["clocks"])
s=d.get("strong_c3") or {}
for k,v in s.items():
    if isinstance(v, dict) and "ms_per_step" in v: print("strong_c3", k, "%.1f ms"%v["ms_per_step"], "reduce %.3f ms"%v.get("reduce_ms_per_step",0))
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or sharding or page_locked" > gpurun_out/r02_final_pytest_n$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02_final_pytest_n$N.log
