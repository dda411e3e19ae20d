"""One line per call: median rollout-launch time of the swapped kernel on C2 (m = 1) + the phase trace of CTA 0.
usage: python tools/tcs_one.py [label] [debug bits]"""
import os
import sys
sys.path.insert(0, ".")
label = sys.argv[1] if len(sys.argv) > 1 else "tcs"
bits = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
sys.argv = [sys.argv[0], "C2", "1"]
src = open("tools/tcs_sweep.py").read().split("\nrun(1, 0, 4)")[0]
ns = {"__name__": "tcs_one"}
exec(compile(src, "tools/tcs_sweep.py", "exec"), ns)
print("==", label, flush=True)
ns["run"](2, 32, 4, trace=True, skew=bits << 20)
