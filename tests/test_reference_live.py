"""Live parity sweep against the reference's own planner graph code (build container only).

tests/golden/live_reference_check.py draws cases from a seed -- all six environments, CEM / random shooting / discrete,
E in {1,2,3,5}, 1-3 particles per member, 50-80 candidates, 1-4 environments, with and without the context encoder,
probabilistic and deterministic -- and runs each through the UNMODIFIED builders of /root/reference (over the NumPy
TensorFlow stand-in) and through the oracle.  Skipped where /root/reference does not exist (the GPU box): the committed
recordings (tests/test_reference_pinned.py) are the travelling form of the same check."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference/cadm/dynamics/core/utils.py"


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_equals_the_reference_graph_on_a_seeded_sweep(seed):
    count = 40
    run = subprocess.run([sys.executable, os.path.join(HERE, "golden", "live_reference_check.py"), str(seed), str(count)],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    rows = [json.loads(line) for line in run.stdout.splitlines() if line.startswith("{")]
    assert len(rows) == count
    for r in rows:
        assert r["returns_rel"] <= 1e-12 and r["plan_abs"] <= 1e-12 and r["index_equal"], r
    specs = [r["spec"] for r in rows]
    # the sweep reaches every branch the planner has
    assert {s["mode"] for s in specs} == {"cem", "rs", "rs_discrete"}
    assert {s["envname"] for s in specs} == {"halfcheetah", "cripple_halfcheetah", "ant", "slim_humanoid", "pendulum", "cartpole"}
    assert any(s["context"] and s["m"] > 1 and s["E"] > 1 and s["mode"] == "cem" for s in specs)      # quirks Q2 / Q3
    assert any(s["det"] for s in specs) and any(not s["det"] and s["p"] > s["E"] for s in specs)      # Q1 with several particles
    assert sum(r["spread"] > 1e-3 for r in rows) > count // 2                                          # candidates are separated
