import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytrace_clj_b200 as rt
from scripts.tc_sweep import run
for name in ("c2", "c4"):
    for ctas in (0, 140, 132, 120, 104, 88):
        run(name, {"cull_tc": 1, "tc_ctas": ctas})
