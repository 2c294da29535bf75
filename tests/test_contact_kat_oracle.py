"""Analytic contact KATs (tests/contact_kats.py) on the CPU oracle -- the physics pins of the oracle's otherwise "parity unpinned"
contact model.  The CUDA path is held to the same answers in tests/test_gpu_contact_kat.py."""
import numpy as np

from tests.contact_kats import check, run_all
from tests.oracle_util import OracleEnv


class OracleRunner:
    def __init__(self, task):
        self.env = OracleEnv(task, "joints")

    def set(self, row):
        self.env.set_full_state(row)

    def step(self, action):
        self.env.step(np.asarray(action, np.float32))
        return self.env.full_state()

    def close(self):
        self.env.close()


def test_contact_kats_on_the_oracle():
    res = run_all(OracleRunner)
    for k, v in res.items():
        print(f"{k}: got {v[0]:.6f} want {v[1]:.6f} tol {v[2]:.2e}")
    check(res)
