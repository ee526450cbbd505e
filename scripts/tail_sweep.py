"""wf_tail's warp-per-path threshold (option tail_solo) on C2 / C1 / C4: python scripts/tail_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.tc_sweep import run

if __name__ == "__main__":
    for name in ("c2", "c1", "c4"):
        for solo in (0, 8, 16, 24, 32, 48, 64, 128):
            run(name, {"tail_solo": solo}, reps=8)
