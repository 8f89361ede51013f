"""Developer tool (needs a GPU and the oracle): a random campaign for the any-collider path -- k_step_small
(speculate + commit, block-wide interval search) and, with PG_GENERAL_ROW_SWEEP=1, the row-sweep kernel --
against the oracle: random mixed worlds (planes, static AABBs, dynamic AABBs / spheres / items, capsule player or
none), random sizes, restitutions and time steps, stepped in random chunk lengths with scene nodes refreshed.
Usage: python tests/random_campaign_mixed.py [cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np

import gen
import ora
from crux_b200 import gpu, scenes


def run(cases: int, seed: int, steps: int = 120, verbose: bool = True):
    orc = ora.OracleLib()
    rng = np.random.default_rng(seed)
    total_events = 0
    for case in range(cases):
        with_player = bool(rng.integers(2))
        n_dyn = int(rng.choice([3, 8, 16, 24, 40]))
        n_box = int(rng.choice([0, 2, 6]))
        L = float(rng.choice([3.0, 6.0, 9.0]))
        dt = np.float32(rng.choice([1.0 / 60.0, 1.0 / 60.0, 0.004, 0.02]))
        st, dy, pl = gen.mixed_world(rng, gpu.node_transform, n_static_boxes=n_box, n_dyn=n_dyn, with_player=with_player, L=L)
        g = gpu.GpuWorlds(1, len(st), len(dy), max(len(pl), 1), max_events_per_world=8192)
        g.set_bodies(0, scenes.CLASS_STATIC, st)
        g.set_bodies(0, scenes.CLASS_DYNAMIC, dy)
        if len(pl):
            g.set_bodies(0, scenes.CLASS_PLAYER, pl)
        o = ora.World(orc, st, dy, pl)
        done = 0
        while done < steps:
            k = int(rng.integers(1, 17))
            g.step(dt, k, refresh_nodes=True)
            o.step(dt, k, refresh_nodes=1)
            done += k
            for cls in (0, 1, 2):
                a, b = g.get_bodies(0, cls), o.read(cls)
                assert gen.state_equal(a, b), (case, done, cls)
                assert gen.same_bits(a["node_xform"], b["node_xform"]), (case, done, cls)
            ge, oe = g.events(), o.events()
            key = lambda e: [(int(x["step"]), int(x["event_type"]), int(x["a_class"]), int(x["a_index"]), int(x["b_class"]), int(x["b_index"])) for x in e]
            assert key(ge) == key(oe), (case, done, len(ge), len(oe))
            total_events += len(ge)
        g.close(); o.close()
    if verbose:
        print(f"{cases} random mixed worlds x {steps} steps: {total_events} contact events, state / node matrices / event order bit-exact")
    return total_events


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
