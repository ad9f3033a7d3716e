"""One-shot all-reduce of a data-parallel step's two small exchanges over NVLink peer memory (marf_peer_allreduce,
csrc/peer_allreduce.cu) instead of NCCL: the ranks' gradient buffers, loss sums and flag arrays live in ONE symmetric
allocation per rank (torch.distributed._symmetric_memory), whose device pointers are exchanged once at construction."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


class PeerAllReduce:
    def __init__(self, device, n_grad: int, group=None):
        import torch.distributed._symmetric_memory as symm
        self.lib = L.load()
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise RuntimeError("marf_peer_allreduce serves up to 8 ranks of one box")
        # layout in 4-byte words: [grads fp32 | sums fp64 x 8 | flags uint32 x (2 world + 1)], every part 16-byte aligned
        a4 = lambda x: (x + 3) // 4 * 4
        self.n_grad = n_grad
        self.off_sums = a4(n_grad)
        self.off_flags = self.off_sums + 16
        words = self.off_flags + a4(2 * self.world + 1)
        try:
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:                      # (newer torch enables it implicitly)
            pass
        self.buf = symm.empty(words, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        if len(base) != self.world or base[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned unexpected buffer pointers")
        vp = C.c_void_p * self.world
        self._in_grad = vp(*base)
        self._in_sums = vp(*[b + 4 * self.off_sums for b in base])
        self._flags = vp(*[b + 4 * self.off_flags for b in base])
        self.grad_local = self.buf[:n_grad]
        self.sums = self.buf[self.off_sums:self.off_sums + 16].view(torch.float64)      # [8] fp64, reduced in place
        self.seq = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)               # every rank's flags are zero before the first round

    def _call(self, dtype, ins, out_ptr, n):
        self.seq += 1
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.marf_peer_allreduce(self.device.index or 0, dtype, ins, self._flags, self.rank, self.world,
                                          C.c_void_p(out_ptr), n, self.seq & 0xFFFFFFFF, st)
        if rc:
            raise L.MarfError(f"marf_peer_allreduce failed (code {rc})")

    def allreduce_sums(self):
        """self.sums <- sum over the ranks, in place."""
        self._call(1, self._in_sums, self.sums.data_ptr(), 8)
        return self.sums

    def allreduce_grads(self, out: torch.Tensor):
        """out[:n_grad] <- sum over the ranks of grad_local (out is a plain fp32 tensor of this rank)."""
        assert out.dtype == torch.float32 and out.numel() >= self.n_grad and out.is_contiguous()
        self._call(0, self._in_grad, out.data_ptr(), self.n_grad)
        return out
