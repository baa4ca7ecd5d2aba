"""Multi-GPU plumbing (one process per GPU, torch.distributed): replicated dense parameters are
all-reduced, embedding tables are row-sharded (owner = id % G, local row = id // G)."""
import torch
import torch.distributed as dist


class DenseGradAllReduce(object):
    """Sum-reduce the gradients of the replicated parameters in one flat bucket and average them
    (data-parallel mean loss).  NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests."""

    def __init__(self, module, skip=()):
        self.params = [p for n, p in module.named_parameters() if p.requires_grad and n not in skip]

    def reduce(self):
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
