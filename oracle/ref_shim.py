"""TEST INFRASTRUCTURE — not part of the product path.

Loads the UNMODIFIED reference (alexfrom0815/Online-3D-BPP-PCT, mounted read-only at
/root/reference) inside this container so that golden vectors can be recorded from it and
the C restatement in oracle/pct_oracle.c can be pinned against it.

The reference needs three interpreter-level accommodations;
no reference file is edited or copied:
  1. `gym` is not installed            -> a stub module exposing only what the reference touches
  2. `np.float` was removed in numpy>=1.24 (pct_envs/*/convex_hull.py:42) -> alias to float
  3. sys.path gets /root/reference

Only `tests/` (when /root/reference exists), `tests/golden/make_golden.py` and the survey-style
CPU timing scripts may import this module.  Nothing here runs on the GPU box.
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PCT_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pct_envs"))


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")

    class Env(object):
        action_space = None
        observation_space = None
        metadata = {}
        spec = None
        reward_range = (-float("inf"), float("inf"))

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env=None):
            self.env = env
            self.action_space = getattr(env, "action_space", None)
            self.observation_space = getattr(env, "observation_space", None)
            self.metadata = getattr(env, "metadata", {})

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kw):
            return self.env.reset(**kw)

        def close(self):
            return self.env.close()

    class ObservationWrapper(Wrapper):
        pass

    class RewardWrapper(Wrapper):
        pass

    spaces = types.ModuleType("gym.spaces")

    class Box(object):
        def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
            self.low, self.high = low, high
            self.shape = tuple(shape) if shape is not None else ()
            self.dtype = np.dtype(dtype)

    class Dict(dict):
        pass

    class Tuple(tuple):
        pass

    spaces.Box, spaces.Dict, spaces.Tuple = Box, Dict, Tuple
    spaces.box = types.ModuleType("gym.spaces.box")
    spaces.box.Box = Box

    registry = {}
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")

    def register(id, entry_point=None, **kw):
        registry[id] = entry_point

    def make(id, **kwargs):
        import importlib
        mod, cls = registry[id].split(":")
        return getattr(importlib.import_module(mod), cls)(**kwargs)

    registration.register = register
    envs.registration = registration
    core = types.ModuleType("gym.core")
    core.Wrapper, core.Env = Wrapper, Env
    gym.Env, gym.Wrapper = Env, Wrapper
    gym.ObservationWrapper, gym.RewardWrapper = ObservationWrapper, RewardWrapper
    gym.spaces, gym.envs, gym.core = spaces, envs, core
    gym.make, gym.register = make, register
    gym._registry = registry
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.spaces.box": spaces.box, "gym.envs": envs,
                        "gym.envs.registration": registration, "gym.core": core})


def load_reference():
    """Make `pct_envs`, `wrapper`, `tools`, ... importable from the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    if not hasattr(np, "float"):
        np.float = float  # noqa: removed alias used by convex_hull.py:42 (do NOT touch np.bool)
    _install_gym_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # np.cross on 2-vectors is deprecated in numpy 2 (convex_hull.py:104) – keep the log readable
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import pct_envs.PctDiscrete0 as D  # noqa
    import pct_envs.PctContinuous0 as C  # noqa
    return D, C


def make_stream_creator(module, items):
    """An item source with exactly RandomBoxCreator's draw discipline
    (pct_envs/*/binCreator.py:24-39: one draw at reset, one after every successful placement)
    but yielding a caller-supplied sequence instead of np.random.randint."""
    base = module.binCreator.BoxCreator

    class StreamCreator(base):
        def __init__(self, seq):
            super().__init__()
            self.seq = [tuple(s) for s in seq]
            self.pos = 0

        def generate_box_size(self, **kw):
            self.box_list.append(self.seq[self.pos % len(self.seq)])
            self.pos += 1

    return StreamCreator(items)


def load_policy_module():
    """The reference's policy network (model.DRL_GAT over attention_model.AttentionModel), unmodified except for the ONE in-memory
    line Python >= 3.8 needs (`super()` inside a NamedTuple body raises at class creation;
    attention_model.py:25 becomes `return tuple.__getitem__(self, key)`).  No reference file is edited or copied."""
    import importlib
    load_reference()
    if "attention_model" not in sys.modules:
        path = os.path.join(REFERENCE_ROOT, "attention_model.py")
        src = open(path).read().replace("return super(AttentionModelFixed, self).__getitem__(key)", "return tuple.__getitem__(self, key)")
        mod = types.ModuleType("attention_model")
        mod.__file__ = path
        argv, sys.argv = sys.argv, sys.argv[:1]  # tools.py parses sys.argv in get_args(); importing it must not see pytest's flags
        try:
            sys.modules["attention_model"] = mod
            exec(compile(src, path, "exec"), mod.__dict__)
        except Exception:
            sys.modules.pop("attention_model", None)
            raise
        finally:
            sys.argv = argv
    return importlib.import_module("model"), importlib.import_module("tools")
