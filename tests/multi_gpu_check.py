"""Multi-GPU parity check (run under torchrun on N GPUs of one box; not collected by pytest):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank trains one step of the row-sharded TEM on its own batch; rank-local results are compared
with an UNsharded model (same parameters) that processes every rank's batch and averages: the loss of
the rank's batch, the replicated dense gradients, the rows of the item-table gradient this rank owns,
and the sharded full-catalog top-k against the unsharded top-k."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from golden_util import DEFAULTS
    from prodsearch_b200 import synth
    from prodsearch_b200.item_transformer import ItemTransformerRanker, ShardedItemTransformerRanker
    cfg = dict(DEFAULTS)
    cfg.update(embedding_size=128, ff_size=512, heads=8, inter_layers=1, neg_per_pos=5, dropout=0.0)
    cfg = argparse.Namespace(**cfg)
    P, V, B = 40000, 5000, 96
    torch.manual_seed(1)
    ref = ItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V))
    torch.manual_seed(1)
    shd = ShardedItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V))
    ref.train(), shd.train()
    batches = [synth.tem_batch(B, P, V, seed=50 + r) for r in range(world)]

    def dev(b):
        return argparse.Namespace(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(b).items()})
    b, ni, nw = batches[rank]
    shd.injected_negatives = (ni.cuda(), nw.cuda())
    loss = shd(dev(b))
    shd.zero_grad()
    loss.backward()
    shd.sync_grads()
    total = None
    ref.zero_grad()
    for r in range(world):
        br, nir, nwr = batches[r]
        ref.injected_negatives = (nir.cuda(), nwr.cuda())
        l = ref(dev(br))
        if r == rank:
            assert abs(float(l) - float(loss)) <= 1e-5 * abs(float(l)), (float(l), float(loss))
        total = l if total is None else total + l
    (total / world).backward()
    worst = 0.0
    for (k, p), (k2, p2) in zip(ref.named_parameters(), shd.named_parameters()):
        assert k == k2
        g_ref = p.grad if p.grad is not None else torch.zeros_like(p)
        g = p2.grad if p2.grad is not None else torch.zeros_like(p2)
        if k == "product_emb.weight":
            g_ref = g_ref[rank::world]
        scale = float(g_ref.abs().max()) + 1e-12
        err = float((g - g_ref).abs().max())
        assert err <= 1e-4 * scale + 2e-7, (k, err, scale)
        worst = max(worst, err / scale)
    # sharded catalog ranking
    q = torch.randn(7 + rank, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9 + rank))
    ids, sc = shd.rank_catalog(q, k=100)
    ids_r, sc_r = ref.rank_catalog(q, k=100)
    assert torch.equal(ids, ids_r) and torch.equal(sc, sc_r)
    dist.barrier()
    if rank == 0:
        print("multi_gpu_check ok: world=%d, worst relative gradient error %.2e, sharded top-100 == unsharded" % (world, worst))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
