"""wf_tail's lane-group-per-path phase (options tail_solo, tail_lpp) on C2 / C1 / C4: python scripts/tail_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.tc_sweep import run

if __name__ == "__main__":
    for name in ("c2", "c1", "c4"):
        run(name, {"tail_solo": 0}, reps=8)
        for lpp, solos in ((32, (24,)), (16, (24, 48, 96)), (8, (32, 64, 96, 128, 192, 256, 512))):
            for solo in solos:
                run(name, {"tail_solo": solo, "tail_lpp": lpp}, reps=8)
