"""One collision scenario, real nocturne_cpp (oracle/_ref) vs the C restatement with contacts (oracle/sim_oracle.c),
compared bit for bit every step.  Run as a script in a fresh process: python tests/contact_case.py {rear,side,head,pile}
(used by tests/test_oracle.py::test_contact_solver_bit_exact_vs_reference_nocturne)."""
import json
import math
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def collision_scene(mode):
    from ctrlsim_b200.synth import make_scene
    sc = make_scene(5, n_vehicles=6, n_roads=6, n_chunks=4, lane_ids=[3])
    o0, o1, o2 = sc["json"]["objects"][:3]
    x0, y0, h0 = o0["position"][0]["x"], o0["position"][0]["y"], o0["heading"][0]
    c, s = math.cos(math.radians(h0)), math.sin(math.radians(h0))

    def place(o, x, y, hdeg, speed):
        for k in range(len(o["position"])):
            o["position"][k] = {"x": x, "y": y}
            o["heading"][k] = hdeg
            o["velocity"][k] = {"x": speed * math.cos(math.radians(hdeg)), "y": speed * math.sin(math.radians(hdeg))}

    if mode == "rear":      # a fast car runs into a slow one from behind, slightly offset and yawed
        place(o0, x0, y0, h0, 2.0)
        place(o1, x0 - 9.0 * c, y0 - 9.0 * s + 0.4, h0 + 3.0, 9.0)
    elif mode == "side":    # T-bone
        place(o0, x0, y0, h0, 5.0)
        place(o1, x0 + 6 * c - 7 * s, y0 + 6 * s + 7 * c, h0 - 90, 6.0)
    elif mode == "head":    # head-on, offset
        place(o0, x0, y0, h0, 6.0)
        place(o1, x0 + 14 * c, y0 + 14 * s + 0.8, h0 + 180 + 5, 5.0)
    elif mode == "pile":    # three cars in a row: the last pushes the middle one into the first (one island, two contacts)
        place(o0, x0, y0, h0, 0.5)
        place(o1, x0 - 6.5 * c, y0 - 6.5 * s, h0, 4.0)
        place(o2, x0 - 13.5 * c, y0 - 13.5 * s + 0.3, h0 + 2.0, 10.0)
    else:
        raise ValueError(mode)
    return sc


def dense(sid=2):
    """BASELINE config-2 scene (64 vehicles), uniformly random controls for everyone, vehicles vanishing from step 60:
    up to 15 simultaneous touching contacts in islands of several vehicles.  Vehicles parked at (-1e6, -1e6) are not
    compared: the reference solves the contacts among them, the restatement skips them (never observed)."""
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.synth import make_scene
    from oracle import ref_shims, sim_port
    ref_shims.install()
    import nocturne
    cfg = default_config()
    sc = make_scene(sid)
    path = tempfile.mktemp(suffix=".json")
    with open(path, "w") as f:
        json.dump(sc["json"], f)
    sim = nocturne.Simulation(scenario_path=path, config=cfg.nocturne["scenario"])
    vehs = sim.getScenario().vehicles()
    for v in vehs:
        v.expert_control = False
        v.physics_simulated = True
    port = sim_port.ScenePort(sim_port.parse_scenario(sc["json"]), contacts=True)
    rng = np.random.default_rng(sid)
    touched = 0
    for t in range(91):
        p = np.array([[v.getPosition().x, v.getPosition().y] for v in vehs])
        live = p[:, 0] > -500000
        assert (p[live] == port.position()[live]).all(), ("dense", t, np.abs(p - port.position())[live].max())
        assert (np.array([v.getHeading() for v in vehs])[live] == port.heading()[live]).all(), ("dense", t)
        assert (np.array([v.getSpeed() for v in vehs])[live] == port.speed()[live]).all(), ("dense", t)
        touched = max(touched, port.n_touching())
        for i, v in enumerate(vehs):
            a, s = rng.uniform(-3, 3), rng.uniform(-0.3, 0.3)
            if t >= 60 and i % 7 == 3:
                v.setPosition(-1000000, -1000000)
                port.teleport(i, -1000000, -1000000)
            if a > 0:
                v.acceleration = a
            else:
                v.brake(abs(a))
            v.steering = s
            port.set_action(i, a, s)
        sim.step(0.1)
        port.step(0.1)
    assert touched >= 8, touched
    os.remove(path)
    print(f"dense: bit-exact for 90 steps, up to {touched} touching contacts at once")


def main(mode):
    if mode == "dense":
        return dense()
    from ctrlsim_b200.config import default_config
    from oracle import ref_shims, sim_port
    ref_shims.install()
    import nocturne
    cfg = default_config()
    sc = collision_scene(mode)
    path = tempfile.mktemp(suffix=".json")
    with open(path, "w") as f:
        json.dump(sc["json"], f)
    sim = nocturne.Simulation(scenario_path=path, config=cfg.nocturne["scenario"])
    vehs = sim.getScenario().vehicles()
    for v in vehs:
        v.expert_control = False
        v.physics_simulated = True
    port = sim_port.ScenePort(sim_port.parse_scenario(sc["json"]), contacts=True)
    rng = np.random.default_rng(1)
    touched, hit = 0, 0
    for t in range(45):
        p = np.array([[v.getPosition().x, v.getPosition().y] for v in vehs])
        assert (p == port.position()).all(), (mode, t, np.abs(p - port.position()).max())
        assert (np.array([v.getHeading() for v in vehs]) == port.heading()).all(), (mode, t)
        assert (np.array([v.getSpeed() for v in vehs]) == port.speed()).all(), (mode, t)
        assert (np.array([int(v.collision_type_veh) != 0 for v in vehs]) == port.collisions()[0]).all(), (mode, t)
        touched = max(touched, port.n_touching())
        hit += int(port.collisions()[0].any())
        for i, v in enumerate(vehs):
            a, s = (0.5, 0.0) if i < 3 else (rng.uniform(-2, 2), rng.uniform(-0.2, 0.2))
            if t == 30 and i == 1:
                v.setPosition(-1000000, -1000000)  # a vehicle that runs out of logged states mid-contact (evaluator.py:176)
                port.teleport(i, -1000000, -1000000)
            if a > 0:
                v.acceleration = a
            else:
                v.brake(abs(a))
            v.steering = s
            port.set_action(i, a, s)
        sim.step(0.1)
        port.step(0.1)
    assert touched >= (2 if mode == "pile" else 1) and hit >= 2, (mode, touched, hit)
    os.remove(path)
    print(f"{mode}: bit-exact for 45 steps, up to {touched} touching contacts, {hit} steps with a collision flag")


if __name__ == "__main__":
    main(sys.argv[1])
