"""AddressSanitizer / UBSan run of the PRODUCT's contact code on the host (tests/host_contacts_shim.cpp builds
ctrlsim_b200/csrc/sim_contacts.cuh for the CPU): collision cases plus two 64-vehicle scenes with random controls and
vehicles teleported away.  Usage: bash tools/contacts_asan.sh"""
import sys, ctypes, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from contact_case import collision_scene
from oracle import sim_port
from ctrlsim_b200.synth import make_scene
H = ctypes.CDLL(os.environ.get('HC_SO', '/tmp/libhc_asan.so'))
F, U8 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
H.hc_init.argtypes = [F, F, F, ctypes.c_int, ctypes.c_int, F]
H.hc_world_step.argtypes = [F, F, F, ctypes.c_int, ctypes.c_int, F, U8, ctypes.c_float]
fields = ("px", "py", "cx", "cy", "lcx", "lcy", "ang", "vx", "vy", "om", "sleep_t", "thr", "brk", "steer", "awake")
fp = lambda a: a.ctypes.data_as(F)
L = sim_port.lib()
def run(parsed, steps, rng, teleport_at=None):
    B = sim_port.ScenePort(parsed, contacts=False); n = N = parsed["n"]
    pack = lambda: np.ascontiguousarray(np.stack([B.arr[k].astype(np.float32) for k in fields] + [np.zeros(n, np.float32)]))
    cstate = np.zeros(H.hc_words(N), np.float32); body = pack()
    H.hc_init(fp(body), fp(B.arr["len"]), fp(B.arr["wid"]), N, n, fp(cstate))
    for t in range(steps):
        tele = np.zeros(n, np.uint8)
        for i in range(n):
            if teleport_at is not None and t >= teleport_at and i % 5 == 1:
                B.teleport(i, -1000000, -1000000); tele[i] = 1
            B.set_action(i, rng.uniform(-3, 3), rng.uniform(-0.3, 0.3))
        L.simo_freecar_all(ctypes.byref(B.s), np.float32(0.1))
        body = pack()
        H.hc_world_step(fp(body), fp(B.arr["len"]), fp(B.arr["wid"]), N, n, fp(cstate), tele.ctypes.data_as(U8), np.float32(0.1))
        for k, row in zip(fields, body): B.arr[k][:] = row.astype(B.arr[k].dtype)
        L.simo_finish_step(ctypes.byref(B.s), sim_port._fp(B.segs), len(B.segs))
rng = np.random.default_rng(0)
for mode in ("rear", "side", "head", "pile"):
    run(sim_port.parse_scenario(collision_scene(mode)["json"]), 45, rng, teleport_at=30)
for sid in (2, 5):  # 64-vehicle config-2 scenes with random controls: dozens of pairs, up to 16 touching contacts
    run(sim_port.parse_scenario(make_scene(sid)["json"]), 90, rng, teleport_at=60)
print("asan/ubsan drive finished")
