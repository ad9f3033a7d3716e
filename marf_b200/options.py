"""Configuration surface of the drop-in: `--key.sub=value` flags over `options/<yaml>.yaml`.

Mirrors the behaviour of the reference's options.py (parse_arguments :14-39, set_opt :42-56, load_options :59-73,
override_options :76-96, process_options :99-120, save_options_file :123-150) with the same function names, so
`train.py --model=planar --yaml=planar --barf_c2f=[0,0.4] --arch.posenc!` means the same thing.  Differences, all
additive: new keys (`precision`, `synthetic`, `max_chunk_pixels`) have defaults in planar.yaml; prompts are skipped
when stdin is not a TTY (unknown keys are then rejected / existing option files overwritten); the device is chosen
per rank from LOCAL_RANK under torchrun.
"""
import os
import random
import string
import sys

import numpy as np
import torch
import yaml

from .attrdict import AttrDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_arguments(args):
    """--a.b=c -> {a:{b:c}} (value parsed as YAML); --a.b= -> None; --a.b -> True; --a.b! -> False."""
    tree = {}
    for arg in args:
        if not arg.startswith("--"):
            raise ValueError(f"options must start with '--': {arg}")
        body = arg[2:]
        if "=" in body:
            key, text = body.split("=", 1)
        elif body.endswith("!"):
            key, text = body[:-1], "false"
        else:
            key, text = body, "true"
        node = tree
        *parents, leaf = key.split(".")
        for p in parents:
            node = node.setdefault(p, {})
        if leaf in node:
            raise ValueError(f"option given twice: {key}")
        node[leaf] = yaml.safe_load(text)
    return AttrDict(tree)


def _yaml_path(name):
    for base in (os.getcwd(), ROOT):
        p = os.path.join(base, name)
        if os.path.isfile(p):
            return p
    raise FileNotFoundError(name)


def load_options(fname):
    """YAML file -> AttrDict, with `_parent_` inheritance (parents are bases, the child overrides)."""
    with open(_yaml_path(fname), encoding="utf-8") as f:
        opt = AttrDict(yaml.safe_load(f))
    parents = opt.pop("_parent_", None)
    if parents:
        for parent in ([parents] if isinstance(parents, str) else parents):
            opt = override_options(load_options(parent), opt, key_stack=[])
    print(f"loading {fname}...")
    return opt


def override_options(opt, opt_over, key_stack=None, safe_check=False):
    """Recursive merge of `opt_over` into `opt`.  With safe_check, keys unknown to the YAML need confirmation."""
    key_stack = key_stack or []
    for key, value in opt_over.items():
        if isinstance(value, dict):
            opt[key] = override_options(opt.get(key, AttrDict()), value, key_stack + [key], safe_check)
            continue
        if safe_check and key not in opt:
            dotted = ".".join(key_stack + [key])
            if sys.stdin is None or not sys.stdin.isatty():
                raise KeyError(f'"{dotted}" not found in original opt (non-interactive run: refusing to add it)')
            answer = None
            while answer not in ("y", "n"):
                answer = input(f'"{dotted}" not found in original opt, add? (y/n) ')
            if answer == "n":
                print("safe exiting...")
                sys.exit()
        opt[key] = value
    return opt


def process_options(opt):
    """Seeding, run name, output path, device."""
    if opt.seed is not None:
        random.seed(opt.seed)
        np.random.seed(opt.seed)
        torch.manual_seed(opt.seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(opt.seed)
        if opt.seed != 0:
            opt.name = f"{opt.name}_seed{opt.seed}"
    else:
        opt.name = f"{opt.name}_" + "".join(random.choice(string.ascii_uppercase) for _ in range(4))
    opt.output_path = f"{opt.output_root}/{opt.group}/{opt.name}"
    os.makedirs(opt.output_path, exist_ok=True)
    assert isinstance(opt.gpu, int)
    # data-parallel launch (torchrun): one process per GPU, device from LOCAL_RANK
    opt.world_size = int(os.environ.get("WORLD_SIZE", "1"))
    opt.rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(opt.gpu)))
    if opt.cpu:
        raise RuntimeError("--cpu: marf_b200 has no CPU path (the reference's --cpu is itself unsupported, "
                           "options/planar.yaml:30); use the oracle under oracle/ for CPU checks")
    if not torch.cuda.is_available():
        raise RuntimeError("marf_b200 needs a CUDA device (B200); no CPU fallback exists")
    opt.device = f"cuda:{local if opt.world_size > 1 else opt.gpu}"


def set_opt(opt_cmd=None):
    opt_cmd = opt_cmd or AttrDict()
    print("setting configurations...")
    assert "model" in opt_cmd and "yaml" in opt_cmd, "--model and --yaml are required"
    opt = override_options(load_options(f"options/{opt_cmd.yaml}.yaml"), opt_cmd, key_stack=[], safe_check=True)
    process_options(opt)
    _print_options(opt)
    return opt


def _print_options(opt, level=0):
    for key, value in sorted(opt.items()):
        if isinstance(value, dict):
            print("   " * level + f"* {key}:")
            _print_options(value, level + 1)
        else:
            print("   " * level + f"* {key}: {value}")


def save_options_file(opt):
    """Dump the resolved options next to the outputs; ask before replacing a different older dump."""
    fname = f"{opt.output_path}/options.yaml"
    current = opt.to_dict()
    if os.path.isfile(fname):
        with open(fname, encoding="utf-8") as f:
            old = yaml.safe_load(f)
        if old == current:
            print("existing options file found (identical)")
        else:
            print("existing options file found (different from current one)...")
            if sys.stdin is not None and sys.stdin.isatty():
                answer = None
                while answer not in ("y", "n"):
                    answer = input("override? (y/n) ")
                if answer == "n":
                    print("safe exiting...")
                    sys.exit()
    else:
        print("(creating new options file...)")
    if int(os.environ.get("RANK", "0")) == 0:
        with open(fname, "w", encoding="utf-8") as f:
            yaml.safe_dump(current, f, default_flow_style=False, indent=4)
