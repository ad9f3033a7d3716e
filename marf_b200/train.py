"""Entry point of the drop-in:  python train.py --model=planar --yaml=planar [--barf_c2f=[0,0.4]] [--key=value ...]

Same call sequence as the reference's train.py:11-31 (parse -> options -> import model.<name> -> Model(opt) ->
load_dataset / build_networks / setup_optimizer / setup_visualizer / train).  Under torchrun (WORLD_SIZE>1) one
process per GPU joins an NCCL group first; the flags do not change.
"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

from . import options


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    print(f"Process ID: {os.getpid()}")
    opt = options.set_opt(options.parse_arguments(argv))
    options.save_options_file(opt)
    if opt.world_size > 1 and not dist.is_initialized():
        torch.cuda.set_device(torch.device(opt.device))
        dist.init_process_group("nccl", device_id=torch.device(opt.device))
    with torch.cuda.device(opt.device):
        model = importlib.import_module(f"model.{opt.model}")
        m = model.Model(opt)
        m.load_dataset()
        m.build_networks()
        m.setup_optimizer()
        m.setup_visualizer()
        m.train()
    if dist.is_initialized():
        dist.destroy_process_group()
    return m


if __name__ == "__main__":
    main()
