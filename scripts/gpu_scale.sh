# usage: bash scripts/gpu_scale.sh N   — the bench at N GPUs (weak C2, strong C3), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, then bench args
  name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  echo "$name exit $?"; tail -c 400 gpurun_out/$name.err
  python - gpurun_out/$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms %.3f"%d["ms_per_step"], "value %.1fM samples/s"%(d["value"]/1e6), "tests %.3fT/s"%(d["tests_per_sec"]/1e12), "frac %.4f"%d["roofline"]["frac"], "e2e %.1fM"%(d["e2e"]["value"]/1e6), d["clocks"])
except Exception as e:
    print("no json:", e)
PY
}
run bench_c2_n$N --steps 50 --warmup 3
run bench_c3_strong_n$N --workload c3 --scaling strong --steps 2 --warmup 3 --no-cpu-baseline
if [ "$N" != "1" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or sharding" > gpurun_out/pytest_n$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_n$N.log
fi
