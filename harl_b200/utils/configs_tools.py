"""Config loading / CLI overrides / run-directory plumbing.

Same call contract as the reference's harl/utils/configs_tools.py (get_defaults_yaml_args :9,
update_args :29, get_task_name :48, init_dir :72, save_config :129): three plain dicts
(main args, algo args, env args), yaml defaults per algo and env, ``--key value`` overrides
matched against any same-named leaf, and a results directory
``<log_dir>/<env>/<task>/<algo>/<exp>/seed-XXXXX-<time>/{logs,models}`` with ``config.json``.
"""
import json
import time
from pathlib import Path

import yaml

CONFIG_ROOT = Path(__file__).resolve().parent.parent / "configs"


def _load_yaml(path):
    if not path.exists():
        raise FileNotFoundError(f"no default config {path.name} under {path.parent}")
    with path.open(encoding="utf-8") as fh:
        return yaml.safe_load(fh) or {}


def get_defaults_yaml_args(algo, env):
    """(algo_args, env_args) from configs/algos_cfgs/<algo>.yaml and configs/envs_cfgs/<env>.yaml."""
    return (_load_yaml(CONFIG_ROOT / "algos_cfgs" / f"{algo}.yaml"),
            _load_yaml(CONFIG_ROOT / "envs_cfgs" / f"{env}.yaml"))


def update_args(unparsed_dict, *args):
    """Overwrite every leaf whose key appears in ``unparsed_dict`` (at any nesting depth), in place."""
    def walk(node):
        for key, val in node.items():
            if isinstance(val, dict):
                walk(val)
            elif key in unparsed_dict:
                node[key] = unparsed_dict[key]

    for cfg in args:
        walk(cfg)


_TASK_KEY = {
    "smac": lambda e: e["map_name"],
    "smacv2": lambda e: e["map_name"],
    "mamujoco": lambda e: f"{e['scenario']}-{e['agent_conf']}",
    "pettingzoo_mpe": lambda e: f"{e['scenario']}-{'continuous' if e['continuous_actions'] else 'discrete'}",
    "gym": lambda e: e["scenario"],
    "football": lambda e: e["env_name"],
    "dexhands": lambda e: e["task"],
    "lag": lambda e: f"{e['scenario']}-{e['task']}",
    "synthetic": lambda e: e.get("task", "synthetic"),
}


def get_task_name(env, env_args):
    return _TASK_KEY[env](env_args)


class NullWriter:
    """Stand-in when no TensorBoard writer can be imported (metrics still go to stdout / progress.txt)."""

    def add_scalars(self, *a, **k):
        pass

    def export_scalars_to_json(self, *a, **k):
        pass

    def close(self):
        pass


def _make_writer(log_path):
    try:
        from tensorboardX import SummaryWriter  # what the reference uses (configs_tools.py:86)
        return SummaryWriter(str(log_path))
    except ImportError:
        pass
    try:
        from torch.utils.tensorboard import SummaryWriter

        w = SummaryWriter(str(log_path))
        if not hasattr(w, "export_scalars_to_json"):
            w.export_scalars_to_json = lambda *a, **k: None
        return w
    except Exception:  # tensorboard missing or broken: logging must never stop training
        return NullWriter()


def init_dir(env, env_args, algo, exp_name, seed, logger_path):
    """Create the run directory tree; returns (run_dir, log_dir, models_dir, writer)."""
    stamp = time.strftime("%Y-%m-%d-%H-%M-%S", time.localtime())
    run = Path(logger_path) / env / get_task_name(env, env_args) / algo / exp_name / f"seed-{seed:0>5}-{stamp}"
    logs, models = run / "logs", run / "models"
    logs.mkdir(parents=True, exist_ok=True)
    models.mkdir(parents=True, exist_ok=True)
    return str(run), str(logs), str(models), _make_writer(logs)


def convert_json(obj):
    """Best-effort JSON-serialisable copy of a config tree."""
    if isinstance(obj, dict):
        return {str(k): convert_json(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [convert_json(v) for v in obj]
    try:
        json.dumps(obj)
        return obj
    except TypeError:
        return getattr(obj, "__name__", str(obj))


def save_config(args, algo_args, env_args, run_dir):
    """Dump {main_args, algo_args, env_args} to <run_dir>/config.json (reloadable with --load_config)."""
    blob = convert_json({"main_args": args, "algo_args": algo_args, "env_args": env_args})
    with open(Path(run_dir) / "config.json", "w", encoding="utf-8") as fh:
        fh.write(json.dumps(blob, separators=(",", ":\t"), indent=4, sort_keys=True))
