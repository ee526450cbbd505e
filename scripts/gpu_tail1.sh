mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "tail_warp_per_path or trace_paths_replay or cornell_paths_replay or identical_paths" > gpurun_out/r3_t1.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r3_t1.log
timeout 300 python scripts/tail_sweep.py > gpurun_out/tail_sweep_4.txt 2>&1; cat gpurun_out/tail_sweep_4.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader
