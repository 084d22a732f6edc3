"""Minimal stand-in for `gymnasium` — TEST INFRASTRUCTURE ONLY.

The reference (Zhehui-Huang/quad-swarm-rl) imports gymnasium at module scope
(quadrotor_dynamics.py:4, quadrotor_control.py:2, quadrotor_single.py:22,
quadrotor_multi.py:6) but the hot path only needs `spaces.Box`, `Env`, `Wrapper`
and `utils.seeding.np_random`.  gymnasium is not installed in this image and
there is no network, so the oracle harness puts this directory on sys.path
before importing the unmodified reference.  Nothing under quad_swarm_rl_b200/
imports it.
"""
from . import spaces  # noqa: F401
from . import utils  # noqa: F401
from . import error  # noqa: F401


class Env:
    metadata = {}
    render_mode = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)
