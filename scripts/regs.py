"""Registers / spills per kernel from the last nvcc build log (ptxas -v)."""
import re
import subprocess
import sys

log = open(sys.argv[1] if len(sys.argv) > 1 else "raytrace_clj_b200/csrc/build.log").read()
cur = sp = None
for line in log.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:80]
        continue
    if "spill" in line and cur:
        sp = line.strip()
    if "Used" in line and cur:
        print(f"{cur:82s} {re.search(r'Used (\d+) registers', line).group(1):>4s} regs | {sp}")
        cur = None
