import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cfg():
    from ctrlsim_b200.config import default_config
    return default_config()


def load_golden(name):
    import json
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", f"rollout_{name}.npz"))
    spec = json.loads(bytes(g["spec_json"]).decode())
    metrics = json.loads(bytes(g["metrics_json"]).decode())
    return g, spec, metrics


def load_planner_adversary_golden(name):
    """tests/golden/planner_adversary_<name>.npz (oracle/make_golden_planner_adversary.py): per-scene records of the
    unmodified reference PlannerAdversaryEvaluator, its metrics and the generator arguments."""
    import json
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", f"planner_adversary_{name}.npz"))
    spec = json.loads(bytes(g["spec_json"]).decode())
    metrics = json.loads(bytes(g["metrics_json"]).decode())
    recs = []
    for k in range(len(spec["scenes"])):
        pre = f"s{k}_"
        recs.append({key[len(pre):]: g[key] for key in g.files if key.startswith(pre)})
    return recs, spec, metrics
