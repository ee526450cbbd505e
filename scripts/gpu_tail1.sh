mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "tail_warp_per_path or trace_paths_replay or cornell_paths_replay or identical_paths" > gpurun_out/r3_t1.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r3_t1.log
timeout 300 python scripts/tail_sweep.py > gpurun_out/tail_sweep_3.txt 2>&1; cat gpurun_out/tail_sweep_3.txt
