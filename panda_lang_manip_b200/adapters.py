"""Vector-env adapters around ``PandaVecEnv``: the callers on the learner side of the step path.

The reference is used through ``gym.make(id)`` handed to a learner (examples/train_push.py:1-12: stable-baselines3 ``DDPG`` with
``HerReplayBuffer``; examples/reach.py / rgb_rendering.py: the plain Gymnasium loop).  A learner that wants thousands of envs would
wrap N such envs in ``gymnasium.vector.AsyncVectorEnv`` or SB3's ``SubprocVecEnv``; these two classes offer the same interfaces on
top of one batched device handle instead of N processes.  Neither gymnasium nor stable-baselines3 is imported (they are not
installed here): the classes are duck-typed to the methods those libraries' training loops call.

* ``PandaGymVectorEnv``  -- ``gymnasium.vector.VectorEnv`` of gymnasium 0.27-0.28 (the reference's pin, env.yml:58): ``reset(seed, options)``,
  ``step(actions) -> (obs, rewards, terminations, truncations, infos)`` with same-step auto-reset and ``infos["final_observation"]`` /
  ``infos["_final_observation"]``, ``call(name, ...)``, ``single_observation_space`` / ``single_action_space``.
* ``PandaSB3VecEnv``     -- ``stable_baselines3.common.vec_env.VecEnv``: ``reset() -> obs``, ``step_async`` / ``step_wait`` ->
  ``(obs, rewards, dones, infos)`` with ``infos[i]["terminal_observation"]``, ``"TimeLimit.truncated"``, ``"is_success"``;
  ``env_method("compute_reward", ag, dg, infos, indices=...)`` as HerReplayBuffer calls it; ``get_attr`` / ``set_attr`` / ``seed``.

``output="torch"`` keeps every array on the GPU (a torch learner never touches the host); ``output="numpy"`` is the drop-in form.
"""
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .panda_gym import spaces
from .vec_env import PandaVecEnv, compute_reward as _compute_reward


def _batch_box(box, n: int):
    return spaces.Box(np.broadcast_to(box.low, (n,) + box.shape).copy(), np.broadcast_to(box.high, (n,) + box.shape).copy(), dtype=np.float32)


class _Base:
    def __init__(self, task: str, num_envs: int, reward_type: str = "sparse", control_type: str = "ee", device: int = 0, seed: int = 0,
                 output: str = "numpy", env_id_offset: int = 0) -> None:
        if output not in ("numpy", "torch"):
            raise ValueError("output must be 'numpy' or 'torch'")
        self._args = dict(task=task, num_envs=int(num_envs), reward_type=reward_type, control_type=control_type, device=device, env_id_offset=env_id_offset)
        self.output = output
        self.env = PandaVecEnv(seed=seed, auto_reset=False, **self._args)      # the adapter resets finished envs itself (it needs the terminal observation)
        self.num_envs = int(num_envs)
        e = self.env
        # core.py:218-224 (Dict of Box(-10, 10)) and panda.py:33 (Box(-1, 1))
        self.single_observation_space = spaces.Dict(dict(observation=spaces.Box(-10.0, 10.0, shape=(e.obs_dim,), dtype=np.float32),
                                                         achieved_goal=spaces.Box(-10.0, 10.0, shape=(e.goal_dim,), dtype=np.float32),
                                                         desired_goal=spaces.Box(-10.0, 10.0, shape=(e.goal_dim,), dtype=np.float32)))
        self.single_action_space = spaces.Box(-1.0, 1.0, shape=(e.action_dim,), dtype=np.float32)

    # -- helpers -----------------------------------------------------------------------------------------------------
    def _out(self, t: torch.Tensor):
        return t.clone() if self.output == "torch" else t.cpu().numpy()

    def _obs(self, d: Dict[str, torch.Tensor]):
        return {k: self._out(v) for k, v in d.items()}

    def _reseed(self, seed: Optional[int]) -> None:
        if seed is not None:
            self.env.close()
            self.env = PandaVecEnv(seed=int(seed), auto_reset=False, **self._args)

    def _advance(self, actions):
        """One step plus same-step reset of finished envs.  Returns (obs dict of device tensors (reset rows replaced), reward,
        terminated, truncated, done mask, terminal observation dict (device, all rows; valid where done))."""
        e = self.env
        a = torch.as_tensor(actions, dtype=torch.float32, device=e.device)
        obs, rew, term, trunc, _ = e.step(a)
        rew, term, trunc = rew.clone(), term.bool(), trunc.bool()
        done = term | trunc
        final = None
        if bool(done.any()):
            final = {k: v.clone() for k, v in obs.items()}
            obs = e.reset(mask=done)
        return obs, rew, term, trunc, done, final

    def compute_reward(self, achieved_goal, desired_goal, info: Any = None):
        """Task.compute_reward (tasks/*.py) on arbitrary batches; numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
        if torch.is_tensor(achieved_goal) and achieved_goal.is_cuda:
            return _compute_reward(self.env.task, self.env.reward_type, achieved_goal, desired_goal)
        a = torch.as_tensor(np.asarray(achieved_goal), device=self.env.device)
        d = torch.as_tensor(np.asarray(desired_goal), device=self.env.device)
        return _compute_reward(self.env.task, self.env.reward_type, a, d).cpu().numpy()

    def close(self) -> None:
        self.env.close()


class PandaGymVectorEnv(_Base):
    """``gymnasium.vector.VectorEnv`` interface over one batched handle (see module docstring)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.observation_space = spaces.Dict({k: _batch_box(s, self.num_envs) for k, s in self.single_observation_space.spaces.items()})
        self.action_space = _batch_box(self.single_action_space, self.num_envs)
        self.is_vector_env = True
        self.closed = False

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        self._reseed(seed)
        return self._obs(self.env.reset()), {}

    def step(self, actions):
        obs, rew, term, trunc, done, final = self._advance(actions)
        infos: Dict[str, Any] = {"is_success": self._out(term)}
        if final is not None:
            if self.output == "torch":
                infos["final_observation"] = final
            else:       # gymnasium's layout: an object array with one obs dict per finished env, None elsewhere
                fo = np.full(self.num_envs, None, dtype=object)
                host = {k: v.cpu().numpy() for k, v in final.items()}
                for i in np.flatnonzero(done.cpu().numpy()):
                    fo[i] = {k: v[i] for k, v in host.items()}
                infos["final_observation"] = fo
            infos["_final_observation"] = self._out(done)
        return self._obs(obs), self._out(rew), self._out(term), self._out(trunc), infos

    def call(self, name: str, *args, **kwargs):
        attr = getattr(self, name)
        return attr(*args, **kwargs) if callable(attr) else attr

    def close(self, **kwargs) -> None:
        if not self.closed:
            super().close()
            self.closed = True


class PandaSB3VecEnv(_Base):
    """``stable_baselines3.common.vec_env.VecEnv`` interface over one batched handle (see module docstring)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.observation_space = self.single_observation_space
        self.action_space = self.single_action_space
        self.render_mode = None
        self._actions = None
        self.reward_type = self.env.reward_type

    def reset(self):
        return self._obs(self.env.reset())

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        self._reseed(seed)
        return [None if seed is None else int(seed) + i for i in range(self.num_envs)]

    def step_async(self, actions) -> None:
        self._actions = actions

    def step_wait(self):
        obs, rew, term, trunc, done, final = self._advance(self._actions)
        t, u, dn = term.cpu().numpy(), trunc.cpu().numpy(), done.cpu().numpy()
        infos: List[Dict[str, Any]] = [{"is_success": bool(t[i]), "TimeLimit.truncated": bool(u[i] and not t[i])} for i in range(self.num_envs)]
        if final is not None:
            host = {k: v.cpu().numpy() for k, v in final.items()}
            for i in np.flatnonzero(dn):
                infos[i]["terminal_observation"] = {k: v[i] for k, v in host.items()}
        return self._obs(obs), self._out(rew), self._out(done), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        return [indices] if isinstance(indices, int) else list(indices)

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        """HerReplayBuffer calls ``env_method("compute_reward", next_achieved_goal, new_goals, infos, indices=[0])`` and takes [0]."""
        out = getattr(self, method_name)(*method_args, **method_kwargs)
        return [out for _ in self._indices(indices)]

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        return [getattr(self, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False for _ in self._indices(indices)]

    def get_images(self):
        raise NotImplementedError("rendering is out of scope of the B200 step path")
