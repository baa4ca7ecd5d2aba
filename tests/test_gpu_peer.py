"""Peer-memory sharding (psb_peer_*; prodsearch_b200/peer.py) on ONE GPU: the kernels only see pointer arrays,
so G ranks are simulated inside this process (PeerGroup.simulate) and compared with the unsharded answer.
The real multi-process path (CUDA IPC over NVLink) is exercised by tests/multi_gpu_check.py under torchrun."""
import argparse
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _groups(world):
    from prodsearch_b200 import peer
    return peer.PeerGroup.simulate(world, "cuda")


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_peer_gather_is_bit_exact(world):
    from prodsearch_b200 import peer
    torch.manual_seed(0)
    rows, d = 1003, 128
    full = torch.randn(rows, d, device="cuda")
    full[rows - 1] = 0
    groups = _groups(world)
    tables = [peer.PeerShardedTable(rows, d, g, pad_idx=rows - 1, full=full) for g in groups]
    for r, t in enumerate(tables):
        assert torch.equal(t.weight.detach(), full[r::world])
    a = torch.randint(0, rows, (37,), device="cuda")
    b = torch.randint(0, rows, (5, 9), device="cuda")
    b[1, 4:] = rows - 1
    b[3, :] = rows - 1
    for t in tables:
        mini, (ra, rb), pad = t.fetch([a, b])
        assert pad == a.numel() + b.numel() and mini.shape == (pad + 1, d)
        assert torch.equal(mini[ra], full[a]) and torch.equal(mini[rb], full[b])
        assert torch.equal(rb == pad, b == rows - 1) and not bool((ra == pad).any())
        assert torch.equal(mini[pad], full[rows - 1])


def test_peer_gather_flags_bad_index():
    from prodsearch_b200 import _lib, peer
    g, = _groups(1)
    t = peer.PeerShardedTable(10, 8, g, pad_idx=9, full=torch.ones(10, 8, device="cuda"))
    ids = torch.tensor([1, 12, -3, 4], device="cuda")
    out = torch.full((4, 8), 7.0, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().psb_peer_gather_rows(t.shard.ptr_array(), 1, 10, 8, ids.data_ptr(), 4, out.data_ptr(), None,
                                                -1, 0, err.data_ptr(), _lib.stream_ptr()), "gather")
    assert int(err) == 1 and torch.equal(out[1], torch.zeros(8, device="cuda")) and float(out[0].sum()) == 8


@pytest.mark.parametrize("world", [2, 4])
def test_peer_fold_matches_index_add(world):
    """Every simulated rank reduces its own contributions by global id into its staging list; every owner
    folds all lists.  Reference: index_add of all contributions on the full table, then the owner's rows."""
    from prodsearch_b200 import ops, peer
    torch.manual_seed(1)
    rows, d, n = 517, 64, 300
    full = torch.randn(rows, d, device="cuda")
    groups = _groups(world)
    tables = [peer.PeerShardedTable(rows, d, g, pad_idx=rows - 1, full=full, stage_cap=2 * n) for g in groups]
    ref = torch.zeros(rows, d, device="cuda", dtype=torch.float64)
    for r, t in enumerate(tables):
        idx = torch.randint(0, rows, (n,), device="cuda")
        idx[:20] = 5                      # a hot row every rank touches
        idx[20:25] = rows - 1             # pad id: dropped
        src = torch.randn(n, d, device="cuda")
        ops.scatter_reduce([ops.make_contrib(idx, src)], rows, d, rows - 1, want_rows=True, device="cuda",
                           out_uniq=t.stage_rows, out_nu=t.stage_n, out_red=t.stage_vals)
        keep = idx != rows - 1
        ref.index_add_(0, idx[keep], src[keep].double())
    for r, t in enumerate(tables):
        t.fold(1.0 / world)
        want = (ref[r::world] / world).float()
        assert torch.allclose(t.grad, want, rtol=1e-5, atol=1e-5)
        assert t.weight.grad is t.grad
    # reproducible: folding again gives the identical bits
    g0 = tables[0].grad.clone()
    tables[0].fold(1.0 / world)
    assert torch.equal(g0, tables[0].grad)


def test_peer_allreduce_and_sqnorm():
    from prodsearch_b200 import _lib, peer
    world, n = 4, 1027
    groups = _groups(world)
    bufs = [g.alloc(4 * 1028) for g in groups]
    vals = []
    for r, b in enumerate(bufs):
        v = torch.randn(1028, device="cuda")
        b.view(torch.float32, (1028,)).copy_(v)
        vals.append(v)
    want = sum(v.double() for v in vals)[:n] * 0.25
    for g, b in zip(groups, bufs):
        out = torch.zeros(1028, device="cuda")
        g.allreduce(b, n, out, scale=0.25)
        assert torch.allclose(out[:n].double(), want, rtol=1e-6, atol=1e-6) and float(out[n:].abs().sum()) == 0
    # psb_grad_sqnorm
    a, b = torch.randn(5000, device="cuda"), torch.randn(33, device="cuda")
    arr = (_lib.AdamTensor * 2)(_lib.AdamTensor(None, a.data_ptr(), None, None, a.numel()),
                                _lib.AdamTensor(None, b.data_ptr(), None, None, b.numel()))
    lib = _lib.load()
    wb = int(lib.psb_adam_workspace_bytes(arr, 2))
    ws = torch.empty(wb, dtype=torch.uint8, device="cuda")
    out = torch.zeros(1, device="cuda")
    _lib.check(lib.psb_grad_sqnorm(arr, 2, out.data_ptr(), ws.data_ptr(), wb, _lib.stream_ptr()), "sqnorm")
    want = float((a.double() ** 2).sum() + (b.double() ** 2).sum())
    assert abs(float(out) - want) <= 1e-5 * want


def test_peer_barrier_two_streams_and_timeout():
    """Two simulated ranks on two streams pass three barriers; a rank whose peer never arrives times out
    and raises the error flag instead of hanging."""
    from prodsearch_b200 import _lib, peer
    groups = _groups(2)
    lib = _lib.load()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for _ in range(3):
        for g, s in zip(groups, streams):
            _lib.check(lib.psb_peer_barrier(g.flags.ptr_array(), g.rank, 2, g.epoch.data_ptr(), g.err.data_ptr(),
                                            2_000_000_000, g.wait_cycles.data_ptr(), 0, s.cuda_stream), "barrier")
    torch.cuda.synchronize()
    assert int(groups[0].err) == 0 and int(groups[1].err) == 0
    assert int(groups[0].epoch) == 3 and int(groups[1].epoch) == 3
    g = groups[0]
    _lib.check(lib.psb_peer_barrier(g.flags.ptr_array(), 0, 2, g.epoch.data_ptr(), g.err.data_ptr(), 2_000_000,
                                    None, 0, streams[0].cuda_stream), "barrier")
    torch.cuda.synchronize()
    assert int(g.err) == 2                       # 1 + the rank that never arrived
    with pytest.raises(RuntimeError):
        g.check_errors()


@pytest.mark.parametrize("dropout", [0.0])
def test_peer_sharded_tem_step_matches_unsharded(dropout):
    """G = 2 simulated ranks, each with its own batch: one clipped-Adam step of the peer-sharded TEM equals the
    unsharded model trained on the mean of the two batch losses (loss, shard / dense gradients, updated rows)."""
    from golden_util import DEFAULTS
    from prodsearch_b200 import peer, synth
    from prodsearch_b200.item_transformer import ItemTransformerRanker, PeerShardedItemTransformerRanker
    from prodsearch_b200.optimizers import build_optim
    world = 2
    cfg = dict(DEFAULTS)
    cfg.update(embedding_size=128, ff_size=512, heads=8, inter_layers=1, neg_per_pos=5, dropout=dropout, optim="adam",
               lr=0.01, max_grad_norm=0.05, beta1=0.9, beta2=0.999, decay_method="adam", warmup_steps=8000,
               l2_lambda=0.0, train_from="")
    cfg = argparse.Namespace(**cfg)
    P, V, B = 3001, 2000, 48
    groups = _groups(world)
    torch.manual_seed(3)
    ref = ItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V))
    ref_opt = build_optim(cfg, ref)
    models, opts = [], []
    for g in groups:
        torch.manual_seed(3)
        m = PeerShardedItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V), peer=g)
        models.append(m)
        opts.append(build_optim(cfg, m))
        m.train()
    ref.train()
    batches = [synth.tem_batch(B, P, V, seed=70 + r) for r in range(world)]
    wq = max(b.query_word_idxs.shape[1] for b, _, _ in batches)

    def dev(b):
        out = argparse.Namespace(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(b).items()})
        q = torch.full((B, wq), V - 1, dtype=torch.int64, device="cuda")
        q[:, :out.query_word_idxs.shape[1]] = out.query_word_idxs
        out.query_word_idxs = q
        return out
    losses = []
    for m, (b, ni, nw) in zip(models, batches):
        m.injected_negatives = (ni.cuda(), nw.cuda())
        loss = m(dev(b))
        m.zero_grad()
        loss.backward()
        losses.append(loss)
    for m in models:
        m.sync_stage()
    for m in models:
        m.sync_fold()
    for m, o in zip(models, opts):
        m.sync_norm(o)
    total = None
    ref.zero_grad()
    for r, (b, ni, nw) in enumerate(batches):
        ref.injected_negatives = (ni.cuda(), nw.cuda())
        l = ref(dev(b))
        assert abs(float(l) - float(losses[r])) <= 1e-5 * abs(float(l)), (float(l), float(losses[r]))
        total = l if total is None else total + l
    (total / world).backward()
    refp = dict(ref.named_parameters())
    for r, m in enumerate(models):
        for k, p in m.named_parameters():
            g_ref = refp[k].grad if refp[k].grad is not None else torch.zeros_like(refp[k])
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if k in ("product_emb.weight", "word_embeddings.weight"):
                g_ref = g_ref[r::world]
            scale = float(g_ref.abs().max()) + 1e-12
            err = float((g - g_ref).abs().max())
            assert err <= 1e-4 * scale + 2e-7, (k, err, scale)
    ref_opt.step()
    for o in opts:
        o.step()
    tn = float(ref_opt.optimizer.total_norm)
    for r, (m, o) in enumerate(zip(models, opts)):
        assert abs(float(o.optimizer.total_norm) - tn) <= 1e-4 * tn          # global norm, not the shard's
        for k, p in m.named_parameters():
            w_ref = refp[k].detach()
            if k in ("product_emb.weight", "word_embeddings.weight"):
                w_ref = w_ref[r::world]
            assert torch.allclose(p.detach(), w_ref, rtol=1e-4, atol=2e-6), k
    # the two replicas of the dense parameters are bit-identical (same all-reduce order on every rank)
    for (k, p0), (_, p1) in zip(models[0].named_parameters(), models[1].named_parameters()):
        if k not in ("product_emb.weight", "word_embeddings.weight"):
            assert torch.equal(p0, p1), k
    # evaluation through the shards: candidate scores equal the unsharded model's
    ref.eval()
    tb = dev(batches[0][0])
    tb.candi_prod_idxs = torch.randint(0, P, (B, 17), device="cuda")
    want = ref.test(tb)
    for m in models:
        m.eval()
        assert torch.allclose(m.test(tb), want, rtol=1e-4, atol=1e-5)
