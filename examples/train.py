"""Command-line entry point with the reference's interface (examples/train.py of PKU-MARL/HARL):

    python examples/train.py --algo happo --env pettingzoo_mpe --exp_name test [--key value ...]
    python examples/train.py --load_config path/to/config.json

Unknown ``--key value`` pairs override same-named config leaves; ``--load_config`` reloads a saved or
tuned ``config.json``.  The reference's own examples/train.py also runs unchanged against this repo
(it only needs ``harl.utils.configs_tools`` and ``harl.runners.RUNNER_REGISTRY``, both provided).
"""
import argparse
import ast
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from harl.utils.configs_tools import get_defaults_yaml_args, update_args  # noqa: E402


def parse_value(text):
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def main(argv=None):
    from harl.runners import RUNNER_REGISTRY

    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--algo", type=str, default="happo", choices=sorted(RUNNER_REGISTRY),
                        help="on-policy heterogeneous-agent algorithm")
    parser.add_argument("--env", type=str, default="pettingzoo_mpe", help="environment name")
    parser.add_argument("--exp_name", type=str, default="installtest", help="experiment name")
    parser.add_argument("--load_config", type=str, default="", help="load an existing config.json instead of the yaml defaults")
    known, extra = parser.parse_known_args(argv)
    overrides = {k[2:]: parse_value(v) for k, v in zip(extra[0::2], extra[1::2])}
    args = vars(known)
    if args["load_config"]:
        with open(args["load_config"], encoding="utf-8") as fh:
            blob = json.load(fh)
        args["algo"], args["env"] = blob["main_args"]["algo"], blob["main_args"]["env"]
        algo_args, env_args = blob["algo_args"], blob["env_args"]
    else:
        algo_args, env_args = get_defaults_yaml_args(args["algo"], args["env"])
    update_args(overrides, algo_args, env_args)
    runner = RUNNER_REGISTRY[args["algo"]](args, algo_args, env_args)
    runner.run()
    runner.close()


if __name__ == "__main__":
    main()
