"""Running the reference's own consumers — `model.DRL_GAT`, `evaluation_tools.evaluate` — on a current interpreter and on this env.

The reference (alexfrom0815/Online-3D-BPP-PCT @ 5e088f2) targets Python 3.7 / gym 0.13 / numpy < 1.24 (README.md:52-57).  On a
Python 3.12 / numpy 2 box without gym four things stop it before it ever reaches an environment ("known
blockers"); none of them has to do with the env, and none needs an edit of the reference checkout:

  1. `import gym` (envs.py:4, tools.py:8, evaluation.py:8, pct_envs/*/bin3D.py:4) — `enable()` installs a stub exposing exactly what
     the reference touches (Env, Wrapper, spaces.Box/Dict/Tuple, envs.registration.register, make) when gym is not importable;
  2. `np.float` (pct_envs/*/convex_hull.py:42) — aliased to float (np.bool is left alone: it exists in numpy 2);
  3. attention_model.py:25 uses `super()` inside a NamedTuple body, which Python >= 3.8 rejects at class creation
     — `load_policy_modules()` executes the file's source with that one line replaced in memory (`tuple.__getitem__(self, key)`);
  4. evaluation.py:36-37 is a SyntaxError (missing comma), and evaluation_tools.py:46 `np.save`s a ragged list, which numpy >= 1.24
     refuses — `reference_evaluate()` is evaluation.py:10-56 written out (same calls, same order, the backup() of source files left
     out) around the UNMODIFIED `evaluation_tools.evaluate`, with `np.save` given an object array for ragged input while it runs.

    import pct_b200.compat as compat
    args = compat.reference_args('/path/to/Online-3D-BPP-PCT', ['--evaluate', '--load-dataset', '--dataset-path', 'set.pt', ...])
    compat.reference_evaluate('/path/to/Online-3D-BPP-PCT', args)          # env = pct_b200.PackingDiscrete / PackingContinuous
"""
import contextlib
import importlib
import os
import sys
import time
import types

import numpy as np


def _install_gym_stub():
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

    class Box(object):
        def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
            self.low, self.high = low, high
            self.shape = tuple(shape) if shape is not None else ()
            self.dtype = np.dtype(dtype)

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Dict, spaces.Tuple = Box, type("Dict", (dict,), {}), type("Tuple", (tuple,), {})
    spaces.box = types.ModuleType("gym.spaces.box")
    spaces.box.Box = Box
    registry = {}

    def register(id, entry_point=None, **kw):
        registry[id] = entry_point

    def make(id, **kwargs):
        mod, cls = registry[id].split(":")
        return getattr(importlib.import_module(mod), cls)(**kwargs)

    envs, registration, core = types.ModuleType("gym.envs"), types.ModuleType("gym.envs.registration"), types.ModuleType("gym.core")
    registration.register = register
    envs.registration = registration
    core.Wrapper, core.Env = Wrapper, Env
    gym.Env, gym.Wrapper, gym.ObservationWrapper, gym.RewardWrapper = Env, Wrapper, type("ObservationWrapper", (Wrapper,), {}), type("RewardWrapper", (Wrapper,), {})
    gym.spaces, gym.envs, gym.core, gym.make, gym.register, gym._registry = spaces, envs, core, make, register, registry
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.spaces.box": spaces.box, "gym.envs": envs, "gym.envs.registration": registration,
                        "gym.core": core})


def enable(reference_root):
    """Make the reference's modules importable here: gym stub (only when gym is missing), np.float alias, sys.path."""
    if not os.path.isfile(os.path.join(reference_root, "attention_model.py")):
        raise FileNotFoundError("no Online-3D-BPP-PCT checkout at %r" % (reference_root,))
    try:
        import gym  # noqa: F401
    except ImportError:
        _install_gym_stub()
    if not hasattr(np, "float"):
        np.float = float  # noqa: removed alias used at pct_envs/*/convex_hull.py:42
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


@contextlib.contextmanager
def _argv(argv):
    old, sys.argv = sys.argv, [sys.argv[0]] + list(argv)
    try:
        yield
    finally:
        sys.argv = old


def load_policy_modules(reference_root, strip_asserts=False):
    """-> (model, tools): the reference's `model` (DRL_GAT) and `tools` modules; attention_model.py gets its one-line accommodation in memory.
    strip_asserts: compile attention_model.py like `python -O` does (source text untouched): its forward pass contains
    `assert not torch.isnan(log_p).any()` (attention_model.py:198), a device -> host synchronisation that a CUDA graph cannot capture — needed by
    GraphedRollout(use_graph=True) with the real network; the trainer / evaluator paths keep the asserts."""
    enable(reference_root)
    if strip_asserts and "attention_model" in sys.modules and not getattr(sys.modules["attention_model"], "_pct_b200_optimized", False):
        for m in ("attention_model", "model"):
            sys.modules.pop(m, None)
    if "attention_model" not in sys.modules:
        path = os.path.join(reference_root, "attention_model.py")
        src = open(path).read()
        old = "return super(AttentionModelFixed, self).__getitem__(key)"
        if old not in src:
            raise RuntimeError("attention_model.py:25 is not the line this loader replaces; the checkout is not reference @ 5e088f2")
        mod = types.ModuleType("attention_model")
        mod.__file__ = path
        sys.modules["attention_model"] = mod
        try:
            with _argv([]):
                exec(compile(src.replace(old, "return tuple.__getitem__(self, key)"), path, "exec", optimize=1 if strip_asserts else -1), mod.__dict__)
                mod._pct_b200_optimized = bool(strip_asserts)
        except Exception:
            sys.modules.pop("attention_model", None)
            raise
    with _argv([]):
        return importlib.import_module("model"), importlib.import_module("tools")


def reference_args(reference_root, argv):
    """tools.get_args() (tools.py:107-197) on an explicit argument list instead of the process's own command line."""
    _, tools = load_policy_modules(reference_root)
    with _argv(argv):
        return tools.get_args()


@contextlib.contextmanager
def _ragged_save():
    """evaluation_tools.py:46 `np.save(path, all_episodes)` with episodes of different lengths: numpy < 1.24 built an object array
    silently, numpy >= 1.24 raises.  While the reference's evaluate() runs, ragged input is turned into that object array."""
    orig = np.save

    def save(file, arr, *a, **kw):
        try:
            np.asarray(arr)
        except ValueError:
            obj = np.empty(len(arr), dtype=object)
            for i, v in enumerate(arr):
                obj[i] = v
            arr = obj
            kw.setdefault("allow_pickle", True)
        return orig(file, arr, *a, **kw)

    np.save = save
    try:
        yield
    finally:
        np.save = orig


def reference_evaluate(reference_root, args, env=None, policy=None, custom="pct_b200", work_dir="."):
    """evaluation.py:10-56 (main) around the unmodified evaluation_tools.evaluate; returns the directory holding trajs.npy / result.txt.

    env=None builds the drop-in single-env facade with the kwargs evaluation.py passes to gym.make (:27-40) — except `shuffle`:
    tools.py:112 declares it `type=bool, default=True`, i.e. always True from the command line, and shuffling the leaf rows draws from
    the global numpy RNG (D:bin3D.py:114-115), which has no parity definition here; the facade lists the leaves in EMSPoint order;
    policy=None builds DRL_GAT(args) and, with args.load_model, loads args.model_path (:43-49).  The source backup (:52) is left out."""
    import torch
    model, tools = load_policy_modules(reference_root)
    evaluate = importlib.import_module("evaluation_tools").evaluate
    time_str = custom + "-" + time.strftime("%Y.%m.%d-%H-%M-%S", time.localtime(time.time()))
    device = torch.device("cpu") if getattr(args, "no_cuda", False) else torch.device("cuda", args.device)
    torch.set_num_threads(1)  # evaluation.py:22
    torch.manual_seed(args.seed)
    if device.type == "cuda":
        torch.cuda.set_device(args.device)
        torch.cuda.manual_seed_all(args.seed)
    if env is None:
        import warnings
        from .envs import PackingContinuous, PackingDiscrete
        if getattr(args, "shuffle", False):
            warnings.warn("args.shuffle is set (tools.py:112 makes that the default): leaf rows are NOT shuffled here, they keep the EMSPoint order")
        cls = PackingContinuous if str(args.id).startswith("PctContinuous") else PackingDiscrete
        env = cls(setting=args.setting, container_size=args.container_size, item_set=args.item_size_set, data_name=args.dataset_path,
                  load_test_data=args.load_dataset, internal_node_holder=args.internal_node_holder, leaf_node_holder=args.leaf_node_holder,
                  LNES=args.lnes, shuffle=False, sample_from_distribution=args.sample_from_distribution,
                  sample_left_bound=args.sample_left_bound, sample_right_bound=args.sample_right_bound,
                  device=0 if device.type == "cpu" else args.device)
    if policy is None:
        policy = model.DRL_GAT(args).to(device)
        if getattr(args, "load_model", False):
            policy = tools.load_policy(args.model_path, policy)
    out = os.path.join(os.path.abspath(work_dir), "logs", "evaluation", time_str)
    os.makedirs(out, exist_ok=True)  # backup() creates it in the reference (tools.py:38-47)
    cwd = os.getcwd()
    os.chdir(os.path.abspath(work_dir))  # evaluate() writes to ./logs/evaluation/<timeStr>/ (evaluation_tools.py:46-50)
    try:
        with _ragged_save():
            evaluate(policy, env, time_str, args, device, eval_freq=args.evaluation_episodes, factor=args.normFactor)
    finally:
        os.chdir(cwd)
    return out
