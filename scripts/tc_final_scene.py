import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.tc_sweep import run
run("c2", {"cull_tc": 1}); run("c2", {"cull_tc": 1}); run("c4", {"cull_tc": 1}); run("c5-10k", {"cull_tc": 1}, reps=3); run("final", {"cull_tc": 3}, reps=3)
