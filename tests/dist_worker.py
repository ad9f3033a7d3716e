"""Multi-rank check, run under torchrun on N GPUs:  N-rank gradients / losses == 1-rank gradients / losses.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_worker.py [--precision bf16]
Each rank also evaluates the whole (unsharded) problem with a single-rank engine and compares."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32")
    args = ap.parse_args()
    from marf_b200 import options, planar
    from marf_b200.attrdict import AttrDict
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    rank, world = dist.get_rank(), dist.get_world_size()
    os.chdir(ROOT)
    failures = []
    for case, over in {
        "disk_patches": dict(batch_size=2 * world, use_masks=True, use_implicit_mask=False, use_edges=False),
        "disk_rows": dict(batch_size=3, use_masks=True, use_implicit_mask=False, use_edges=False),
        "nomask_rows": dict(batch_size=3, use_masks=False, use_implicit_mask=False, use_edges=False),
        "implicit_patches_edges": dict(batch_size=world, use_masks=True, use_implicit_mask=True, use_edges=True,
                                       H=360, W=480, patch_H=180, patch_W=240),
        # the edge term needs whole patches: uneven deal (one rank takes two), and more ranks than patches (the last rank idles)
        "disk_edges_uneven": dict(batch_size=world + 1, use_masks=True, use_implicit_mask=False, use_edges=True),
        "implicit_edges_idle_rank": dict(batch_size=world - 1, use_masks=True, use_implicit_mask=True, use_edges=True,
                                         H=360, W=480, patch_H=180, patch_W=240),
    }.items():
        if case.endswith("rows") and 3 % world == 0:
            over["batch_size"] = world + 1
        opt = options.load_options("options/planar.yaml")
        opt.update(model="planar", yaml="planar", H=96, W=128, patch_H=48, patch_W=64, use_homographies=False,
                   precision=args.precision, device=dev, output_path=f"/tmp/marf_dist_{rank}", tb=None, world_size=world, rank=rank)
        opt.update(over)
        opt.synthetic = dict(enabled=True, seed=5, occluders=True)
        torch.manual_seed(3)
        m = planar.Model(opt)
        m.load_dataset()
        m.build_networks()
        g = m.graph
        g.warp_param.weight.data.normal_(0, 0.02)
        g.warp_param.weight.data[0] = 0
        dist.broadcast(g.warp_param.weight.data, 0)
        g.it = 700
        g.neural_image.progress.data.fill_(0.2)
        var = AttrDict(idx=torch.arange(opt.batch_size), images=m.images)
        for _ in range(3):                      # (consecutive rounds of the exchange: flags / sequence numbers must line up)
            g.it = 700
            g.forward(var, mode="train")
        loss_d = [float(x) for x in g.engine.loss_values(g._sums)]
        grad_d = g._grad_flat.clone()
        # the same problem on one rank
        g1 = planar.Graph(opt).to(dev)
        g1.load_state_dict(g.state_dict())
        g1.data_parallel = False
        g1.it = 700
        g1.forward(var, mode="train")
        loss_1 = [float(x) for x in g1.engine.loss_values(g1._sums)]
        grad_1 = g1._grad_flat
        gerr = ((grad_d - grad_1).norm() / (grad_1.norm() + 1e-30)).item()
        lerr = max(abs(a - b) / (abs(b) + 1e-12) for a, b in zip(loss_d, loss_1))
        tol_g, tol_l = (2e-3, 1e-5) if args.precision == "fp32" else (2e-2, 1e-4)
        ok = gerr <= tol_g and lerr <= tol_l
        e = g.engine
        print(f"[rank {rank}/{world}] {case}: shard(batch={0 if e.idle else e.batch}, patch_offset={e.patch_offset}, rows={e.rows}, "
              f"row_offset={e.row_offset}) exchange={'peer' if g._peer_grads else ('peer-sums+nccl' if g._peer is not None else 'nccl')} grad rel-L2 err {gerr:.2e}, "
              f"loss rel err {lerr:.2e} {'OK' if ok else 'FAIL'}", flush=True)
        if not ok:
            failures.append(case)
        g.engine.close()
        g1.engine.close()
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        raise SystemExit(f"rank {rank}: FAILED {failures}")
    if rank == 0:
        print("dist_worker: all cases OK")


if __name__ == "__main__":
    main()
