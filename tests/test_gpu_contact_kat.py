"""Analytic contact KATs (tests/contact_kats.py) on the CUDA path (fp32 product kernels and the fp64 instantiation)."""
import numpy as np
import pytest

from tests.contact_kats import check, run_all

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_contact_kats_on_the_gpu(precision):
    import panda_lang_manip_b200 as p

    class GpuRunner:
        def __init__(self, task):
            self.env = p.PandaVecEnv(task, 1, control_type="joints", precision=precision, auto_reset=False)

        def set(self, row):
            st = self.env.get_state().cpu().numpy()
            st[0, :-1] = row
            self.env.set_state(torch.from_numpy(st))

        def step(self, action):
            self.env.step(torch.from_numpy(np.asarray(action, np.float32)[None]).cuda())
            return self.env.get_state()[0, :-1].cpu().numpy()

        def close(self):
            self.env.close()

    res = run_all(GpuRunner)
    for k, v in res.items():
        print(f"{k}: got {v[0]:.6f} want {v[1]:.6f} tol {v[2]:.2e}")
    check(res)
