mkdir -p gpurun_out
timeout 120 python scripts/tc_trace.py gpurun_out/tail_trace.txt 2 1 > gpurun_out/tail_diag.txt 2>&1
grep -c tail gpurun_out/tail_diag.txt; cat gpurun_out/tail_trace.txt | tail -12
