import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.tc_sweep import run
for name in ("c2", "c4"):
    for per in (0, 96, 128, 192, 256):
        run(name, {"cull_tc": 1, "tc_tiles_per_cta": per})
    run(name, {"cull_tc": 1, "light_block": 64, "tc_tiles_per_cta": 128})
