"""Key figures of one `ncu --set full` capture (.ncu-rep), as text: the metrics north_star asks for
(FP32-pipe utilisation, warp execution efficiency, branch divergence) plus issue, occupancy, DRAM traffic,
then the hot regions of the SASS (sass_hot.py).  usage: summarize_ncu.py capture.ncu-rep > summary.txt"""
import csv, io, os, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(name):
    v, u = m.get(name, ("n/a", ""))
    return f"{v} {u}".strip()
print(f"# {os.path.basename(rep)}: {m['Kernel Name'][0]}  grid {m.get('launch__grid_size', ('?',''))[0]} x block {m.get('launch__block_size', ('?',''))[0]}, "
      f"{m.get('launch__registers_per_thread', ('?',''))[0]} regs/thread")
for label, name in [
    ("duration", "gpu__time_duration.sum"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
    ("FP32 (FMA) pipe utilisation, % of peak while active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("FMA pipe cycles active %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("ALU pipe utilisation %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe utilisation %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("executed IPC (per SM, active)", "sm__inst_executed.avg.per_cycle_active"),
    ("warp execution efficiency: threads per executed instruction (of 32)", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("  not predicated off", "smsp__thread_inst_executed_per_inst_executed.pct"),
    ("branch targets uniform % (branch efficiency)", "smsp__sass_average_branch_targets_threads_uniform.pct"),
    ("divergent branch targets (sum)", "smsp__sass_branch_targets_threads_divergent.sum"),
    ("branch instructions % of executed", "derived__smsp__inst_executed_op_branch_pct"),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("active warps per scheduler", "smsp__warps_active.avg.per_cycle_active"),
    ("executed warp instructions", "smsp__inst_executed.sum"),
    ("dram bytes read", "dram__bytes_read.sum"),
    ("dram bytes written", "dram__bytes_write.sum"),
    ("dram throughput", "dram__bytes.sum.per_second"),
    ("shared-memory bank conflicts (ld / st)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("stall per issue: not selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall per issue: dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("stall per issue: math pipe throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall per issue: long scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall per issue: short scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall per issue: wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall per issue: barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall per issue: no instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
]:
    print(f"{label:72s} {g(name)}")
print("\n# SASS regions (contiguous instructions with equal execution count; share of executed instructions / of stall samples)")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tmp = "/tmp/_src.csv"
open(tmp, "w").write(src)
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "sass_hot.py"), tmp, "0.02"],
                     capture_output=True, text=True).stdout)
