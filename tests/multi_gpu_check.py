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
    ref_grads = reference_grads(ref, batches, dev, world, rank, float(loss))
    worst = 0.0
    for (k, p), (k2, p2) in zip(ref.named_parameters(), shd.named_parameters()):
        assert k == k2
        g_ref = ref_grads[k]
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
    pg = peer_check(rank, world, cfg, ref, batches, dev)
    if pg is not None:
        sparse_peer_check(rank, world, cfg, dev, pg)
    dist.destroy_process_group()


def reference_grads(ref, batches, dev, world, rank, my_loss):
    """Gradients of the UNsharded model for the mean over all ranks' batch losses.  One backward per batch (a
    dense gradient sink holds at most 8 contributions and rewrites its buffer per backward), summed by hand."""
    acc = {k: torch.zeros_like(p) for k, p in ref.named_parameters()}
    for r in range(world):
        br, nir, nwr = batches[r]
        ref.injected_negatives = (nir.cuda(), nwr.cuda())
        ref.zero_grad()
        l = ref(dev(br))
        if r == rank:
            assert abs(float(l) - my_loss) <= 1e-5 * abs(float(l)), (float(l), my_loss)
        (l / world).backward()
        for k, p in ref.named_parameters():
            if p.grad is not None:
                acc[k] += p.grad
    ref.zero_grad()
    return acc


def peer_check(rank, world, cfg, ref, batches, dev):
    """The NVLink peer-memory transport (prodsearch_b200/peer.py): one eager step against the unsharded model,
    then the same step captured and replayed as one CUDA graph per rank."""
    from prodsearch_b200 import peer, synth
    from prodsearch_b200.graph_step import GraphedTrainStep
    from prodsearch_b200.item_transformer import PeerShardedItemTransformerRanker
    from prodsearch_b200.optimizers import build_optim
    pg = peer.try_create()
    if pg is None:
        if rank == 0:
            print("peer_check SKIPPED: CUDA IPC / P2P unavailable on this box")
        return None
    P, V, B = 40000, 5000, 96
    for k_, v_ in dict(optim="adam", lr=0.01, max_grad_norm=0.05, beta1=0.9, beta2=0.999, decay_method="adam",
                       warmup_steps=8000, l2_lambda=0.0, train_from="").items():
        setattr(cfg, k_, v_)
    torch.manual_seed(1)
    m = PeerShardedItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V), peer=pg)
    opt = build_optim(cfg, m)
    ref_opt = build_optim(cfg, ref)
    m.train()
    wq = 12                      # synth.tem_batch draws at most 12 query words

    def padded(b):
        out = dev(b)
        q = torch.full((B, wq), V - 1, dtype=torch.int64, device="cuda")
        q[:, :out.query_word_idxs.shape[1]] = out.query_word_idxs
        out.query_word_idxs = q
        return out
    b, ni, nw = batches[rank]
    m.injected_negatives = (ni.cuda(), nw.cuda())
    loss = m(padded(b))
    m.zero_grad()
    loss.backward()
    m.sync_grads(opt)
    ref_grads = reference_grads(ref, batches, padded, world, rank, float(loss))
    refp = dict(ref.named_parameters())
    sharded_keys = ("product_emb.weight", "word_embeddings.weight")
    worst = 0.0
    for k, p in m.named_parameters():
        g_ref = ref_grads[k]
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        if k in sharded_keys:
            g_ref = g_ref[rank::world]
        scale = float(g_ref.abs().max()) + 1e-12
        err = float((g - g_ref).abs().max())
        assert err <= 1e-4 * scale + 2e-7, (k, err, scale)
        worst = max(worst, err / scale)
    for k, p in ref.named_parameters():       # hand the summed gradients to the reference optimizer
        p.grad = ref_grads[k]
    opt.step()
    ref_opt.step()
    tn = float(ref_opt.optimizer.total_norm)
    assert abs(float(opt.optimizer.total_norm) - tn) <= 1e-4 * tn
    for k, p in m.named_parameters():
        w_ref = refp[k].detach()
        if k in sharded_keys:
            w_ref = w_ref[rank::world]
        assert torch.allclose(p.detach(), w_ref, rtol=1e-4, atol=2e-6), k
    pg.check_errors()
    # sharded catalog ranking through the peer model == unsharded (the tables were just updated identically)
    q = torch.randn(5 + rank, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(19 + rank))
    ids, sc = m.rank_catalog(q, k=100)
    # reference: exact top-k over the table re-assembled from the shards (the shards just took an optimizer step;
    # ranking against the separately updated unsharded copy would compare two slightly different tables)
    from prodsearch_b200 import _lib, ops
    shard = m.item_table.weight.detach()
    rows_max = (P + 1 + world - 1) // world
    pad = torch.zeros(rows_max, shard.shape[1], device="cuda")
    pad[:shard.shape[0]] = shard
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.zeros(P + 1, shard.shape[1], device="cuda")
    for r in range(world):
        n_r = (P + 1 - r + world - 1) // world
        full[r::world] = parts[r][:n_r]
    ids_r, sc_r = ops.catalog_topk(q, full, 100, n_items=P, mode=_lib.TOPK_EXACT)
    assert torch.equal(ids, ids_r) and torch.equal(sc, sc_r)
    # ---- the whole multi-GPU step as one CUDA graph per rank
    del loss                     # drop the eager autograd graphs (their AccumulateGrad nodes live on this stream)
    m.zero_grad()
    ref.zero_grad()
    m.injected_negatives = None
    torch.manual_seed(100 + rank)
    graphed = GraphedTrainStep(m, opt, padded(b), pad_values={"query_word_idxs": V - 1, "u_item_idxs": P},
                               sync_grads=lambda: m.sync_grads(opt))
    losses = []
    for it in range(4):
        bb, _, _ = synth.tem_batch(B, P, V, seed=500 + 10 * it + rank)
        losses.append(graphed(padded(bb)).clone())
    torch.cuda.synchronize()
    pg.check_errors()
    vals = [float(x) for x in losses]
    assert all(v == v and 0 < v < 100 for v in vals), vals
    # replicated parameters stay bit-identical across ranks
    chk = torch.stack([p.detach().double().sum() for k, p in m.named_parameters() if k not in sharded_keys])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(torch.equal(allc[0], c) for c in allc)
    dist.barrier()
    if rank == 0:
        print("peer_check ok: world=%d, worst relative gradient error %.2e, graph-replayed losses %s" %
              (world, worst, ["%.4f" % v for v in vals]))
    return pg


def sparse_peer_check(rank, world, cfg, dev, pg):
    """Row-sharded training with the ROW-SPARSE owner-side Adam (item table: grad_mode="rowsparse"; readers bring
    resting rows up to date on the fly, psb_peer_gather_rows_lazy) against the same sharded model with the dense
    owner-side sweep: same loss at every step (every step reads rows that rested since an earlier one), and after
    the final flush the same item shard up to the few elements whose gradient is rounding noise."""
    from prodsearch_b200 import synth
    from prodsearch_b200.item_transformer import PeerShardedItemTransformerRanker
    from prodsearch_b200.optimizers import build_optim
    P, V, B, steps, lr = int(os.environ.get("PSB_CHECK_ROWS", "40000")), 5000, 96, 7, 0.005
    for k_, v_ in dict(optim="adam", lr=lr, max_grad_norm=5.0, beta1=0.9, beta2=0.999, decay_method="adam",
                       warmup_steps=8000, l2_lambda=0.0, train_from="").items():
        setattr(cfg, k_, v_)
    models, opts = [], []
    for mode in ("dense", "rowsparse"):
        torch.manual_seed(7)
        with torch.device("cuda"):
            m = PeerShardedItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V), peer=pg,
                                                 grad_mode=mode)
        m.train()
        models.append(m)
        opts.append(build_optim(cfg, m))
    assert torch.equal(models[0].item_table.weight, models[1].item_table.weight)

    def padded(b):
        out = dev(b)
        q = torch.full((B, 12), V - 1, dtype=torch.int64, device="cuda")
        q[:, :out.query_word_idxs.shape[1]] = out.query_word_idxs
        out.query_word_idxs = q
        return out
    worst = 0.0
    for s in range(steps):
        b, ni, nw = synth.tem_batch(B, P, V, seed=900 + 17 * s + rank)
        cb = padded(b)
        ls = []
        for m, o in zip(models, opts):
            m.injected_negatives = (ni.cuda(), nw.cuda())
            loss = m(cb)
            m.zero_grad()
            loss.backward()
            m.sync_grads(o)
            o.step()
            ls.append(float(loss.detach()))
            del loss
        rel = abs(ls[0] - ls[1]) / abs(ls[0])
        worst = max(worst, rel)
        assert rel <= 2e-5, (s, ls)
    pg.check_errors()
    assert models[1].item_table.lazy_optim is not None and models[1].item_table.weight.grad is None
    stale = float((models[0].item_table.weight - models[1].item_table.weight).abs().max())
    models[1].eval()                       # flush: every resting row of the shard is brought up to date
    err = (models[0].item_table.weight - models[1].item_table.weight).abs()
    assert float(err.max()) <= 0.05 * lr * steps, float(err.max())
    assert float((err > 3e-6 * steps).float().mean()) <= 2e-3
    dist.barrier()
    if rank == 0:
        print("sparse_peer_check ok: world=%d, %d rows, worst relative loss difference %.1e over %d steps, shard max "
              "difference %.1e before / %.1e after the flush" % (world, P, worst, steps, stale, float(err.max())))


if __name__ == "__main__":
    main()
