"""Acceptance logic shared by the GPU parity test of the free-running rollout (tests/test_gpu_parity.py) and by the CPU
test that runs the same logic on a CPU emulation of the GPU's arithmetic (tests/test_oracle.py)."""
import numpy as np

POS_TOL = 1e-3     # metres, BASELINE.json north_star


def first_contact(g):
    cv = (g["reward"][:, :, 6] * g["existence"]).any(0)
    idx = np.where(cv)[0]
    return int(idx[0]) if len(idx) else 91


def first_flip(tr, g, n, T):
    """First step at which a sampled bin differs from the reference, and the list of differing (t, v, c) RTG draws."""
    rtg = tr["tr_rtg_idx"][0, :n, :T].transpose(1, 0, 2).astype(np.int64)
    act = tr["tr_act_idx"][0, :n, :T].T.astype(np.int64)
    bad_r = np.argwhere(rtg != g["rtg_idx"][:T])
    bad_a = np.argwhere(act != g["act_idx"][:T])
    t_r = int(bad_r[:, 0].min()) if len(bad_r) else T
    t_a = int(bad_a[:, 0].min()) if len(bad_a) else T
    return t_r, t_a, bad_r, rtg



def check_rollout_vs_reference(tr, g, name, trig="glibc", steps=90, min_flip=None):
    """``tr``: trace arrays of ONE scene in SceneBatch.trace() layout; ``g``: the reference fixture.  Returns the first
    step with a marginal draw (90 if none).  ``trig``: "glibc" = the default simulator arithmetic (no difference to the
    reference left: positions / headings compared bit for bit through contacts); "fp64" = CTRLSIM_TRIG=fp64.
    ``steps``: episode length of the fixture; ``min_flip``: earliest step at which a marginal draw is tolerated."""
    n = g["pos"].shape[0]
    tc = first_contact(g)
    t_r, t_a, bad_r, rtg = first_flip(tr, g, n, steps)
    assert t_a >= t_r, (name, "an action draw differs before any RTG draw did", t_a, t_r)
    if t_r < steps:
        assert t_r >= (min(tc + 8, 60) if min_flip is None else min_flip), (name, "draws differ too early", t_r, tc, bad_r[:4].tolist())
        for t, v, c in bad_r[bad_r[:, 0] == t_r]:
            assert abs(int(rtg[t, v, c]) - int(g["rtg_idx"][t, v, c])) == 1, (name, t, v, c)
    T = t_r  # states up to and including step T depend on draws before T only
    assert (tr["tr_exist"][0, :n, :T + 1] == g["existence"][:, :T + 1]).all()
    ctrl = np.stack([g["accel"], g["steer"]], -1)

    def deviations(t0, t1):  # over states t0..t1 and the controls applied at t0..t1-1
        if t1 < t0:
            return (0.0,) * 6
        ex = g["existence"][:, t0:t1 + 1].astype(bool)
        exa = g["existence"][:, t0:t1].astype(bool)
        mx = lambda a, m: float(a[m].max()) if m.any() else 0.0
        return (mx(np.abs(tr["tr_pos"][0, :n, t0:t1 + 1].astype(np.float64) - g["pos"][:, t0:t1 + 1]), ex),
                mx(np.abs(tr["tr_heading"][0, :n, t0:t1 + 1].astype(np.float64) - g["heading"][:, t0:t1 + 1]), ex),
                mx(np.abs(tr["tr_vel"][0, :n, t0:t1 + 1].astype(np.float64) - g["vel"][:, t0:t1 + 1]), ex),
                mx(np.abs(tr["tr_action"][0, :n, t0:t1] - ctrl[:, t0:t1]), exa),
                mx(np.abs(tr["tr_reward"][0, :n, t0:t1 + 1].astype(np.float64) - g["reward"][:, t0:t1 + 1]), ex),
                mx(np.abs(tr["tr_nearest"][0, :n, t0:t1 + 1, 0] - g["nearest_dist"][:, t0:t1 + 1]), ex))

    # up to the first contact the simulator is bit-exact by construction (float tolerances of the trace only)
    t_free = min(T, tc)
    dpos, dhead, dvel, dacc, drew, dnd = deviations(0, t_free)
    assert dpos < POS_TOL and dhead < 1e-5 and dvel < 1e-4 and dacc < 1e-4 and drew < 1e-5 and dnd < 1e-3, \
        (name, "before contact", dpos, dhead, dvel, dacc, drew, dnd)
    dpos, dhead, dvel, dacc, drew, dnd = deviations(t_free, T)
    if trig == "glibc":
        # default mode: the simulator replays the reference's fp32 operation order with glibc's own sinf / cosf / tanf,
        # so states in contact are the reference's bit for bit; the derived quantities keep the pre-contact tolerances
        assert dpos == 0.0 and dhead == 0.0, (name, "in contact (glibc trig must be bit-exact)", dpos, dhead)
        assert dvel < 1e-4 and dacc < 1e-4 and drew < 1e-5 and dnd < 1e-3, (name, "in contact", dvel, dacc, drew, dnd)
    else:
        # CTRLSIM_TRIG=fp64: glibc's sinf / cosf are within 1 ulp but not always correctly rounded, this mode evaluates
        # them in fp64 and rounds once; while vehicles push each other the solver amplifies the difference.  The CPU
        # oracle built with that trig reproduces the GPU numbers exactly (crowded: 0.73 mm, 2.3e-5 rad after 74 steps)
        assert dpos < POS_TOL and dhead < 2e-4 and dvel < 2e-3 and dacc < 5e-2 and drew < 1e-3 and dnd < 5e-3, \
            (name, "in contact", dpos, dhead, dvel, dacc, drew, dnd)
    if tc <= steps:
        assert T > tc, (name, "the comparison must cover the contact phase", T, tc)
    return t_r
