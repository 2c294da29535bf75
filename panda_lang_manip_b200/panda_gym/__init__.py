"""Drop-in for the reference package ``panda_gym`` (reference panda_gym/__init__.py:8-54): the 24 environment ids
Panda{Reach,Push,Slide,PickAndPlace,Stack,Flip}[Joints][Dense]-v3 with max_episode_steps 50 (Stack: 100).

With gymnasium installed the ids are registered there (``gymnasium.make``); ``make`` below works either way and applies the
TimeLimit itself.  ``install()`` aliases this package as ``panda_gym`` so that ``import panda_gym`` picks up the B200 backend.
"""
import sys

__version__ = ""  # the reference's version.txt is empty

ENV_IDS = {}
for _reward in ("sparse", "dense"):
    for _control in ("ee", "joints"):
        for _task, _steps in (("Reach", 50), ("Push", 50), ("Slide", 50), ("PickAndPlace", 50), ("Stack", 100), ("Flip", 50)):
            _id = "Panda{}{}{}-v3".format(_task, "Joints" if _control == "joints" else "", "Dense" if _reward == "dense" else "")
            ENV_IDS[_id] = dict(entry_point="Panda{}Env".format(_task), kwargs={"reward_type": _reward, "control_type": _control}, max_episode_steps=_steps)

try:  # pragma: no cover - gymnasium is optional
    from gymnasium.envs.registration import register as _register
    for _id, _spec in ENV_IDS.items():
        _register(id=_id, entry_point=__name__ + ".envs:" + _spec["entry_point"], kwargs=_spec["kwargs"], max_episode_steps=_spec["max_episode_steps"])
except Exception:
    pass


class TimeLimit:
    """Minimal gymnasium.wrappers.TimeLimit: truncated=True once max_episode_steps steps have elapsed since reset."""

    def __init__(self, env, max_episode_steps: int) -> None:
        self.env, self._max, self._t = env, max_episode_steps, 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kwargs):
        self._t = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._t += 1
        return obs, reward, terminated, truncated or self._t >= self._max, info


def make(env_id: str, **kwargs):
    """gym.make for the 24 ids (TimeLimit applied), independent of gymnasium."""
    from . import envs
    if env_id not in ENV_IDS:
        raise KeyError(f"unknown environment id {env_id!r}")
    spec = ENV_IDS[env_id]
    kw = dict(spec["kwargs"]); kw.update(kwargs)
    return TimeLimit(getattr(envs, spec["entry_point"])(**kw), spec["max_episode_steps"])


def install() -> None:
    """Make ``import panda_gym`` resolve to this package."""
    sys.modules.setdefault("panda_gym", sys.modules[__name__])
