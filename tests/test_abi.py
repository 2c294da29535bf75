"""The C-ABI library loads and exports every symbol include/panda_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "panda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("pg_create", "pg_destroy", "pg_reset", "pg_step", "pg_step_host", "pg_compute_reward", "pg_is_success",
                 "pg_save_state", "pg_restore_state", "pg_remove_state", "pg_get_state", "pg_set_state", "pg_stats", "pg_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from panda_lang_manip_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == _declared()


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import panda_lang_manip_b200 as p
    with pytest.raises(p.PandaB200Error):
        p.PandaVecEnv("reach", 4)
    h = ctypes.c_void_p()
    rc = p.load().pg_create(0, 0, 0, 4, 0, 0, 0, 0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in p.load().pg_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "panda_lang_manip_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "libpanda_oracle" not in src and "oracle_util" not in src and "panda_oracle.h" not in src, os.path.join(d, f)


def test_learner_side_helpers_import_and_sample_on_cpu():
    """adapters import without gymnasium / stable-baselines3 and without a GPU; the HER index sampler is plain torch (CPU here)."""
    import torch
    import panda_lang_manip_b200 as p
    from panda_lang_manip_b200 import adapters
    assert hasattr(adapters, "PandaGymVectorEnv") and hasattr(adapters, "PandaSB3VecEnv")
    gen = torch.Generator().manual_seed(0)
    T, M = 50, 20000
    src = torch.randint(0, 40 * T, (M,), generator=gen)
    start = (src // T) * T
    gi = p.future_goal_indices(start, torch.full_like(src, T), src, her_ratio=0.8, generator=gen)
    rel = gi >= 0
    assert abs(rel.float().mean().item() - 0.8) < 0.02 and bool(((gi >= src) & (gi < start + T))[rel].all())
    with __import__("pytest").raises(p.PandaB200Error):       # no device here: the relabel kernel refuses CPU tensors, there is no fallback
        p.her_relabel("reach", "sparse", torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(2, dtype=torch.long), torch.zeros(2, dtype=torch.long))


def test_ctypes_signatures_match_the_header():
    """Every entry point's ctypes argtypes list has as many entries as the header's declaration has parameters."""
    from panda_lang_manip_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "panda_b200.h")).read(), flags=re.S)
    lib = _lib.load()
    for m in re.finditer(r"\b(pg_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        at = getattr(lib, name).argtypes
        assert (at is None and n == 0) or (at is not None and len(at) == n), (name, n, at)
