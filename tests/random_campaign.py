"""Test infrastructure (not collected by pytest): a long random parity campaign of the sphere kernel against the
oracle.  The suite runs 40 cases (test_gpu_worlds.py); this runs as many as asked on a GPU box:
`python tests/random_campaign.py 800 80`."""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, HERE)
import ora
from crux_b200 import gpu
import test_gpu_worlds as t

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 400
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
if not os.path.exists(ora.ORACLE_SO):
    ora.build_oracles()
orc = ora.OracleLib()
total = 0
for seed in range(2, 2 + (n_cases + 99) // 100):
    total += t._random_world_campaign(gpu, orc, min(100, n_cases), seed, steps)
    print("seed", seed, "ok, contacts so far", total, flush=True)
print("campaign ok:", n_cases, "cases x", steps, "steps,", total, "contacts, all bit-exact")
