"""Writes tests/golden/state_dict_spec.json: the name and shape of every tensor in the UNMODIFIED reference model's
``state_dict()`` (models/ctrl_sim.py:29-30 built from the reference's default YAML sizes). Test infrastructure: the
checkpoint importer (ctrlsim_b200/checkpoint.py) and ``weights.param_spec`` are checked against it on CPU.

Run in the build container (needs /root/reference):  python -m oracle.make_state_dict_spec
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from ctrlsim_b200.config import default_config
    from oracle import ref_shims
    ref_shims.install()
    from models import CtRLSim
    model = CtRLSim(default_config())
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    out = os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")
    with open(out, "w") as f:
        json.dump(spec, f, indent=0, sort_keys=True)
    print(out, len(spec), "tensors", sum(int(__import__("numpy").prod(s)) for s in spec.values()), "parameters")


if __name__ == "__main__":
    main()
