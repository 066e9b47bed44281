"""ORACLE (test infrastructure only): ctypes wrapper around oracle/sim_oracle.c plus the scenario-level glue.

``ScenePort`` mirrors the slice of the pybind11 API the evaluator touches:
``Simulation(scenario_path, config)``, ``veh.acceleration = a`` / ``veh.brake(b)`` / ``veh.steering = s``,
``veh.setPosition``, ``sim.step(dt)``, and the getters - as array operations over all vehicles of one scene.
Scenario loading follows nocturne/cpp/src/scenario.cc:893-1057 (objects valid at t=0 only, heading deg->rad float +
NormalizeAngle, speed = |velocity|, road_edge segments) and the expert replay used for ground truth follows
scenario.cc:276-283 + utils/sim.py:20-65.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_F = ctypes.POINTER(ctypes.c_float)
_U8 = ctypes.POINTER(ctypes.c_uint8)


class _SimO(ctypes.Structure):
    _fields_ = ([("n", ctypes.c_int)] + [(k, _F) for k in ("px", "py", "ang", "vx", "vy", "om", "sleep_t", "cx", "cy", "lcx", "lcy")]
                + [("awake", _U8)] + [(k, _F) for k in ("thr", "brk", "steer", "len", "wid")]
                + [(k, _F) for k in ("ox", "oy", "heading", "speed")] + [("coll_veh", _U8), ("coll_edge", _U8)])


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


_LIB_FP64 = None


_LIB_PORT = None


def lib(fp64_trig=False):
    """The C restatement; ``fp64_trig`` True selects the variant built with the GPU's default trig (fp64, rounded once),
    "glibc_port" the variant built with the product's restatement of glibc's sinf / cosf (see sim_oracle.c)."""
    global _LIB, _LIB_FP64, _LIB_PORT
    if fp64_trig == "glibc_port":
        if _LIB_PORT is None:
            _LIB_PORT = _load("libsim_oracle_glibcport.so")
        return _LIB_PORT
    if fp64_trig:
        if _LIB_FP64 is None:
            _LIB_FP64 = _load("libsim_oracle_fp64trig.so")
        return _LIB_FP64
    if _LIB is None:
        _LIB = _load("libsim_oracle.so")
    return _LIB


def _load(name):
    path = os.path.join(HERE, "_build", name)
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "sim_oracle.c")):
        build()
    L = ctypes.CDLL(path)
    P = ctypes.POINTER(_SimO)
    f = ctypes.c_float
    L.simo_spawn.argtypes = [P, ctypes.c_int, f, f, f, f, f, f]
    L.simo_set_action.argtypes = [P, ctypes.c_int, f, f]
    L.simo_teleport.argtypes = [P, ctypes.c_int, f, f]
    L.simo_step.argtypes = [P, f, _F, ctypes.c_int]
    L.simo_update_collision.argtypes = [P, _F, ctypes.c_int]
    L.simc_create.restype = ctypes.c_void_p
    L.simc_create.argtypes = [ctypes.c_int]
    L.simc_free.argtypes = [ctypes.c_void_p]
    L.simc_init_body.argtypes = [P, ctypes.c_void_p, ctypes.c_int]
    L.simc_teleport.argtypes = [P, ctypes.c_void_p, ctypes.c_int]
    L.simc_step.argtypes = [P, ctypes.c_void_p, f, _F, ctypes.c_int]
    L.simc_num_contacts.argtypes = [ctypes.c_void_p]
    L.simc_num_touching.argtypes = [ctypes.c_void_p]
    L.simo_poly_intersects.argtypes = [ctypes.c_int, _F, _F, ctypes.c_int, _F, _F]
    L.simo_poly_segment_intersects.argtypes = [ctypes.c_int, _F, _F, f, f, f, f]
    L.simo_velocity.argtypes = [P, ctypes.c_int, _F, _F]
    L.simo_freecar_all.argtypes = [P, f]
    L.simo_finish_step.argtypes = [P, _F, ctypes.c_int]
    return L


def _fp(a):
    return a.ctypes.data_as(_F)


def poly_intersects(p1, p2) -> bool:
    a = np.ascontiguousarray(np.asarray(p1, np.float32).T)
    b = np.ascontiguousarray(np.asarray(p2, np.float32).T)
    return bool(lib().simo_poly_intersects(a.shape[1], _fp(a[0]), _fp(a[1]), b.shape[1], _fp(b[0]), _fp(b[1])))


def poly_segment_intersects(p, s0, s1) -> bool:
    a = np.ascontiguousarray(np.asarray(p, np.float32).T)
    return bool(lib().simo_poly_segment_intersects(a.shape[1], _fp(a[0]), _fp(a[1]), s0[0], s0[1], s1[0], s1[1]))


def normalize_angle_f32(deg) -> np.float32:
    """geometry_utils.h:41-58: Radians<float>(d) = d / 180.0 * kPi evaluated in double then narrowed on return (T =
    float), NormalizeAngle<float>: fmod in double against 2*pi, wrap, narrow."""
    rad = np.float32(float(np.float32(deg)) / 180.0 * math.pi)
    ret = np.float32(math.fmod(float(rad), 2.0 * math.pi))  # const T ret with T = float
    r = float(ret)
    if r > math.pi:
        out = r - 2.0 * math.pi
    elif r < -math.pi:
        out = r + 2.0 * math.pi
    else:
        out = r
    return np.float32(out)


def parse_scenario(scen: dict):
    """JSON dict -> arrays (float32 as the reference stores them). Objects invalid at t=0 are dropped
    (scenario.cc:959-961); ids are assigned to the kept ones in order."""
    objs = [o for o in scen["objects"] if bool(o["valid"][0]) and o["type"] == "vehicle"]
    n, T = len(objs), len(objs[0]["position"]) if objs else 0
    pos = np.zeros((n, T, 2), np.float32)
    head = np.zeros((n, T), np.float32)
    speed = np.zeros((n, T), np.float32)
    valid = np.zeros((n, T), bool)
    size = np.zeros((n, 2), np.float32)
    target = np.zeros((n, 4), np.float32)  # x, y, heading, speed of the last valid step
    moving = np.zeros(n, bool)
    for i, o in enumerate(objs):
        size[i] = (o["length"], o["width"])
        gp = o.get("goalPosition", {"x": 0.0, "y": 0.0})
        target[i, :2] = (gp["x"], gp["y"])
        for t in range(T):
            pos[i, t] = (o["position"][t]["x"], o["position"][t]["y"])
            head[i, t] = normalize_angle_f32(o["heading"][t])
            vx, vy = np.float32(o["velocity"][t]["x"]), np.float32(o["velocity"][t]["y"])
            speed[i, t] = np.sqrt(np.float32(vx * vx + vy * vy))  # Vector2D::Norm -> std::sqrt(float)
            valid[i, t] = bool(o["valid"][t])
            if valid[i, t]:
                target[i, 2], target[i, 3] = head[i, t], speed[i, t]
                d = pos[i, t] - target[i, :2]
                dist = np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1]))
                if speed[i, t] > np.float32(0.05) or dist > np.float32(0.2):
                    moving[i] = True
    segs = []
    for road in scen["roads"]:
        g = road["geometry"]
        if road["type"] != "road_edge" or isinstance(g, dict):
            continue
        for k in range(len(g) - 1):
            segs.append((g[k]["x"], g[k]["y"], g[k + 1]["x"], g[k + 1]["y"]))
    segs = np.asarray(segs, np.float32).reshape(-1, 4)
    return dict(n=n, T=T, pos=pos, heading=head, speed=speed, valid=valid, size=size, target=target,
                moving=moving, segs=segs)


class ScenePort:
    """All vehicles of one scene as FreeCars (evaluators/evaluator.py:33-41)."""

    def __init__(self, parsed: dict, contacts: bool = False, fp64_trig: bool = False):
        """``contacts``: run the world step with the Box2D contact restatement (sim_oracle.c, second half) instead of the
        contact-free subset."""
        self.p = parsed
        self.contacts = None
        self.L = lib(fp64_trig)  # fp64_trig: the variant with the GPU's trig rounding (sim_oracle.c SIMO_TRIG_FP64)
        n = parsed["n"]
        self.n = n
        self.arr = {k: np.zeros(n, np.float32) for k in
                    ("px", "py", "ang", "vx", "vy", "om", "sleep_t", "cx", "cy", "lcx", "lcy", "thr", "brk", "steer", "len", "wid", "ox", "oy",
                     "heading", "speed")}
        for k in ("awake", "coll_veh", "coll_edge"):
            self.arr[k] = np.zeros(n, np.uint8)
        self.s = _SimO(n=n)
        for k, a in self.arr.items():
            setattr(self.s, k, a.ctypes.data_as(_U8 if a.dtype == np.uint8 else _F))
        self.segs = np.ascontiguousarray(parsed["segs"], np.float32)
        L = self.L
        for i in range(n):
            L.simo_spawn(ctypes.byref(self.s), i, parsed["pos"][i, 0, 0], parsed["pos"][i, 0, 1],
                         parsed["heading"][i, 0], parsed["speed"][i, 0], parsed["size"][i, 0], parsed["size"][i, 1])
        if contacts:
            self.contacts = L.simc_create(n)
            for i in range(n):
                L.simc_init_body(ctypes.byref(self.s), self.contacts, i)
        L.simo_update_collision(ctypes.byref(self.s), _fp(self.segs), len(self.segs))  # scenario.cc:263

    def __del__(self):
        if getattr(self, "contacts", None) and getattr(self, "L", None) is not None:
            try:
                self.L.simc_free(self.contacts)
            except Exception:  # interpreter shutdown
                pass
            self.contacts = None

    def set_action(self, i, accel, steer):
        self.L.simo_set_action(ctypes.byref(self.s), i, np.float32(accel), np.float32(steer))

    def teleport(self, i, x, y):
        self.L.simo_teleport(ctypes.byref(self.s), i, np.float32(x), np.float32(y))
        if self.contacts:
            self.L.simc_teleport(ctypes.byref(self.s), self.contacts, i)

    def step(self, dt=0.1):
        if self.contacts:
            self.L.simc_step(ctypes.byref(self.s), self.contacts, np.float32(dt), _fp(self.segs), len(self.segs))
        else:
            self.L.simo_step(ctypes.byref(self.s), np.float32(dt), _fp(self.segs), len(self.segs))

    def n_touching(self):
        return int(self.L.simc_num_touching(self.contacts)) if self.contacts else 0

    # getters (float32 values widened to python float, like pybind)
    def position(self):
        return np.stack([self.arr["ox"], self.arr["oy"]], -1).astype(np.float64)

    def heading(self):
        return self.arr["heading"].astype(np.float64)

    def speed(self):
        return self.arr["speed"].astype(np.float64)

    def velocity(self):
        h, s = self.arr["heading"], self.arr["speed"]
        # PolarToVector2D(speed, heading) with std::cos/std::sin(float): evaluate with the same libm via C
        vx = np.zeros(self.n, np.float32)
        vy = np.zeros(self.n, np.float32)
        L = self.L
        for i in range(self.n):
            L.simo_velocity(ctypes.byref(self.s), i, _fp(vx[i:i + 1]), _fp(vy[i:i + 1]))
        return np.stack([vx, vy], -1).astype(np.float64)

    def collisions(self):
        return self.arr["coll_veh"].astype(bool), self.arr["coll_edge"].astype(bool)


def ground_truth(parsed: dict, steps: int = 90):
    """utils/sim.py:20-65 on the expert replay: traj[n, steps+1, 8] =
    (x, y, heading, speed, existence, target_x, target_y, length) as float64 views of float32 values."""
    n = parsed["n"]
    gt = np.zeros((n, steps + 1, 8), np.float64)
    gt[:, :, 0:2] = parsed["pos"][:, :steps + 1]
    gt[:, :, 2] = parsed["heading"][:, :steps + 1]
    gt[:, :, 3] = parsed["speed"][:, :steps + 1]
    gt[:, :, 4] = (parsed["pos"][:, :steps + 1, 0] != np.float32(-10000.0)).astype(np.float64)
    gt[:, :, 5:7] = parsed["target"][:, None, :2]
    gt[:, :, 7] = parsed["size"][:, None, 0]
    return gt
