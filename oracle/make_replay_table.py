"""ORACLE (test infrastructure only): BASELINE config 4 reference table.

    python -m oracle.make_replay_table [--scenes 120]

Log-replays scenes ctrlsim_b200.synth.make_replay_scene(0 .. N-1) through the C restatement of the reference simulator
(oracle/sim_oracle.c: inverse bicycle model -> FreeCar -> Box2D world step incl. contact response -> collision / off-road
flags; bit-identical to the real nocturne_cpp, tests/test_oracle.py) and writes the per-scene figures
(collision / off-road vehicle-steps, ADE, position checksums) to tests/golden/replay_oracle.json.  The GPU test and
tools/replay_eval.py compare the CUDA simulator against this table."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=120)
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.synth import make_replay_scene, replay_scene_summary
    from oracle.policy_port import RolloutPort
    cfg = default_config()
    port = RolloutPort(cfg, model=None, eval_threshold=0)
    rows = []
    t0 = time.time()
    for i in range(args.scenes):
        sc = make_replay_scene(i)
        rec = port.run_scene(i, sc["json"], sc["preproc"], replay_only=True)
        rows.append(replay_scene_summary(rec["pos"], rec["heading"], rec["existence"], rec["reward"], rec["gt_pos"]))
    out = {"what": "log replay of make_replay_scene(i), i < scenes, through oracle/sim_oracle.c (= reference nocturne_cpp)",
           "scenes": args.scenes, "rows": rows}
    with open(os.path.join(ROOT, "tests", "golden", "replay_oracle.json"), "w") as f:
        json.dump(out, f)
    tot = {k: sum(r[k] for r in rows) for k in ("veh_steps", "coll_steps", "off_steps", "coll_veh", "off_veh")}
    print(f"{args.scenes} scenes in {time.time() - t0:.0f}s: {tot}")


if __name__ == "__main__":
    main()
