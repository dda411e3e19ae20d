"""Diagnostic switches of the swapped-operand kernel on the C2 workload (results are garbage with a switch on): launch time
and phase trace with (1) no weight copies, (2) no hidden-layer epilogue work, (3) both -- what bounds a layer."""
import os
import sys
sys.path.insert(0, ".")
sys.argv = [sys.argv[0], "C2", "1"]
import importlib.util
spec = importlib.util.spec_from_file_location("tcs_sweep_mod", "tools/tcs_sweep.py")
src = open("tools/tcs_sweep.py").read().split("\nrun(1, 0, 4)")[0]        # the definitions only
ns = {"__name__": "tcs_diag"}
exec(compile(src, "tools/tcs_sweep.py", "exec"), ns)
for name, bits in (("production kernel (no diagnostics compiled in)", 0), ("no weight copies", 1), ("no hidden epilogue", 2), ("neither", 3)):
    print("==", name, flush=True)
    ns["run"](2, 32, 4, trace=True, skew=bits << 20)
