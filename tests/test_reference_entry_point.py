"""The reference's OWN examples/train.py (unmodified, run from /root/reference where that checkout exists) drives this
package through the `harl` alias: it imports harl.utils.configs_tools and harl.runners.RUNNER_REGISTRY, builds the
configuration from this repo's yaml defaults / a tuned config.json, applies the command-line overrides and constructs
RUNNER_REGISTRY[algo](args, algo_args, env_args) -> run() -> close().  The runner class is replaced by a recorder (a real
run needs a GPU; tests/test_gpu_iteration.py covers it), everything before it is the reference's code."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRAIN = os.path.join(os.environ.get("HARL_REFERENCE", "/root/reference"), "examples", "train.py")

DRIVER = r"""
import json, runpy, sys
sys.path.insert(0, {root!r})          # `harl` resolves to this repo's alias package, not to the reference
import harl, harl.runners
assert harl.__file__.startswith({root!r}), harl.__file__
calls = []
class Recorder:
    def __init__(self, args, algo_args, env_args):
        calls.append(dict(args=args, algo_args=algo_args, env_args=env_args))
    def run(self): calls[-1]["run"] = True
    def close(self): calls[-1]["close"] = True
for k in list(harl.runners.RUNNER_REGISTRY):
    harl.runners.RUNNER_REGISTRY[k] = Recorder
sys.argv = ["train.py"] + {argv!r}
runpy.run_path({train!r}, run_name="__main__")
print("RECORD " + json.dumps(calls))
"""


def _run(argv):
    code = DRIVER.format(root=ROOT, argv=argv, train=REF_TRAIN)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RECORD ")][-1]
    return json.loads(line[len("RECORD "):])


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="needs the reference checkout (build container only)")
def test_reference_train_py_builds_the_runner_from_this_repos_configs():
    calls = _run(["--algo", "happo", "--env", "pettingzoo_mpe", "--exp_name", "alias", "--n_rollout_threads", "4096",
                  "--continuous_actions", "False", "--lr", "0.001", "--hidden_sizes", "[64, 64]"])
    assert len(calls) == 1 and calls[0].get("run") and calls[0].get("close")
    c = calls[0]
    assert c["args"]["algo"] == "happo" and c["args"]["env"] == "pettingzoo_mpe" and c["args"]["exp_name"] == "alias"
    assert c["algo_args"]["train"]["n_rollout_threads"] == 4096          # command-line overrides reached the leaves
    assert c["algo_args"]["model"]["lr"] == 0.001 and c["algo_args"]["model"]["hidden_sizes"] == [64, 64]
    assert c["env_args"]["continuous_actions"] is False
    # the defaults are the reference's own (harl/configs/algos_cfgs/happo.yaml), key for key
    import yaml

    ref_yaml = os.path.join(os.path.dirname(os.path.dirname(REF_TRAIN)), "harl", "configs", "algos_cfgs", "happo.yaml")
    ref = yaml.safe_load(open(ref_yaml))
    for section, leaves in ref.items():
        assert set(leaves) == set(c["algo_args"][section]), section
        for k, v in leaves.items():
            if (section, k) not in (("train", "n_rollout_threads"), ("model", "lr"), ("model", "hidden_sizes")):
                assert c["algo_args"][section][k] == v, (section, k)


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("algo", ["hatrpo", "haa2c", "mappo"])
def test_reference_train_py_load_config_and_other_on_policy_algorithms(algo, tmp_path):
    tuned = os.path.join(os.path.dirname(os.path.dirname(REF_TRAIN)), "tuned_configs", "pettingzoo_mpe", "simple_spread_v2-continuous",
                         algo, "config.json")
    if not os.path.exists(tuned):
        pytest.skip("no tuned config for this algorithm / scenario")
    calls = _run(["--load_config", tuned, "--exp_name", "tuned"])
    blob = json.load(open(tuned))
    c = calls[0]
    assert c["args"]["algo"] == blob["main_args"]["algo"] == algo
    assert c["algo_args"] == blob["algo_args"] and c["env_args"] == blob["env_args"]
