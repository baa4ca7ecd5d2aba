"""GPU: the row-sparse, lazily caught-up Adam (psb_adam_sparse_step / psb_adam_rows_catchup) against
the reference's optimizer step -- torch.optim.Adam(eps=1e-9) + clip_grad_norm_ on DENSE gradients
(models/optimizers.py:186,:205-243) -- over many steps in which rows are touched at different times, so that resting
rows have to be brought up to date (series of the skipped updates + closed-form moment decay)."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dense_reference(R, d, shapes, lr, betas, max_norm, noam, warm):
    g = torch.Generator().manual_seed(1)
    table = torch.nn.Parameter(torch.randn(R, d, generator=g).cuda())
    bias = torch.nn.Parameter(torch.randn(R, generator=g).cuda())
    dense = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    return table, bias, dense


@pytest.mark.parametrize("betas,steps,max_norm,noam", [((0.9, 0.999), 25, 5.0, False), ((0.9, 0.999), 25, 0.05, True),
                                                       ((0.5, 0.999), 140, 5.0, False), ((0.9, 0.999), 12, 0.0, False)])
def test_sparse_adam_is_dense_equivalent(betas, steps, max_norm, noam):
    from prodsearch_b200 import _lib
    from prodsearch_b200 import functional as F_
    from prodsearch_b200.optimizers import Optimizer
    R, d, lr, warm = 3000, 128, 5e-3, 20
    shapes = [(128, 128), (513,)]
    ref_t, ref_b, ref_d = _dense_reference(R, d, shapes, lr, betas, max_norm, noam, warm)
    my_t, my_b = torch.nn.Parameter(ref_t.detach().clone()), torch.nn.Parameter(ref_b.detach().clone())
    my_d = [torch.nn.Parameter(p.detach().clone()) for p in ref_d]
    ref_opt = torch.optim.Adam([ref_t, ref_b] + ref_d, lr=lr, betas=betas, eps=1e-9)
    opt = Optimizer("adam", lr, max_norm, beta1=betas[0], beta2=betas[1], decay_method="noam" if noam else "adam",
                    warmup_steps=warm)
    opt.set_parameters([("t", my_t), ("b", my_b)] + [("d%d" % i, p) for i, p in enumerate(my_d)])
    kmax = _lib.load().psb_adam_catchup_steps(betas[0], betas[1])
    if betas[0] == 0.5:
        assert kmax < 60 < steps                     # this case exercises the closed-form tail of the catch-up
    g = torch.Generator().manual_seed(2)
    pad = R - 1
    for step in range(1, steps + 1):
        # rows of this step: a hot set touched every step, a warm set every 7th, one-off cold rows; never the pad row
        n = 40 + int(torch.randint(0, 30, (1,), generator=g))
        rows = torch.cat([torch.arange(0, 16), torch.arange(100, 140) if step % 7 == 0 else torch.empty(0, dtype=torch.long),
                          torch.randint(200, R - 1, (n,), generator=g)]).unique()
        if step == steps:                              # long-resting rows come back at the very end
            rows = torch.cat([rows, torch.arange(100, 140)]).unique()
        nu = rows.numel()
        cap = nu + 13                                  # lists are longer than their valid prefix
        vals = torch.randn(cap, d, generator=g) * (0.01 if step % 2 else 2.0)
        bvals = torch.randn(cap, generator=g) * 0.1
        # -- "forward": the rows about to be read are made current, and must equal the dense reference's rows
        read = torch.cat([rows, torch.randint(0, R, (50,), generator=g)]).cuda()
        F_.ensure_current(my_t, (read,))
        if step > 1:
            torch.testing.assert_close(my_t.detach()[read], ref_t.detach()[read], rtol=0, atol=2e-6 * step)
            torch.testing.assert_close(my_b.detach()[read], ref_b.detach()[read], rtol=0, atol=2e-6 * step)
        # -- dense reference step
        gt = torch.zeros(R, d)
        gt[rows] = vals[:nu]
        gb = torch.zeros(R)
        gb[rows] = bvals[:nu]
        ref_t.grad, ref_b.grad = gt.cuda(), gb.cuda()
        dg = [torch.randn(s, generator=g) * 0.1 for s in shapes]
        for p, q, x in zip(ref_d, my_d, dg):
            p.grad = x.clone().cuda()
            q.grad = x.clone().cuda()
        if noam:
            ref_opt.param_groups[0]["lr"] = lr * min(step ** -0.5, step * warm ** -1.5)
        if max_norm:
            total = torch.nn.utils.clip_grad_norm_([ref_t, ref_b] + ref_d, max_norm)
        ref_opt.step()
        # -- row-sparse step: (unique rows, reduced rows, device count), as the gradient sinks hand them over
        rows_dev = torch.full((cap,), -7, dtype=torch.int32).cuda()
        rows_dev[:nu] = rows.int().cuda()
        nu_dev = torch.tensor([nu], dtype=torch.int32).cuda()
        my_t.row_grad = (rows_dev, vals.cuda(), nu_dev)
        my_b.row_grad = (rows_dev, bvals.cuda(), nu_dev)
        my_t._psb_row_bias, my_t._psb_drop_idx = my_b, pad
        opt.step()
        assert my_t.row_grad is None
        if max_norm:
            assert abs(float(opt.optimizer.total_norm) - float(total)) <= 1e-5 * float(total)
        for p, q in zip(ref_d, my_d):
            assert (p.detach() - q.detach()).abs().max().item() <= 2e-6 * step
    # stale until flushed: the untouched rows of the table still hold old values
    stale = (my_t.detach() - ref_t.detach()).abs().max().item()
    assert stale > 1e-4
    opt.flush()
    tol = 2e-6 * steps
    assert (my_t.detach() - ref_t.detach()).abs().max().item() <= tol
    assert (my_b.detach() - ref_b.detach()).abs().max().item() <= tol
    sd, ref_sd = opt.optimizer.state_dict(), ref_opt.state_dict()
    assert float(sd["state"][0]["step"]) == steps
    for i in range(2 + len(shapes)):
        torch.testing.assert_close(sd["state"][i]["exp_avg"], ref_sd["state"][i]["exp_avg"], rtol=2e-5, atol=1e-7)
        torch.testing.assert_close(sd["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"], rtol=2e-5, atol=1e-9)
    # flushing twice changes nothing; the pad row never moved
    before = my_t.detach().clone()
    opt.flush()
    assert torch.equal(before, my_t.detach())


def test_sparse_adam_rejects_weight_decay_and_dense_grads_on_lazy_tables():
    from prodsearch_b200.optimizers import Optimizer
    t = torch.nn.Parameter(torch.randn(64, 32).cuda())
    opt = Optimizer("adam", 1e-3, 5.0, weight_decay=0.01)
    opt.set_parameters([("t", t)])
    t.row_grad = (torch.arange(4, dtype=torch.int32).cuda(), torch.randn(4, 32).cuda(), torch.tensor([4], dtype=torch.int32).cuda())
    with pytest.raises(RuntimeError):
        opt.step()
    opt = Optimizer("adam", 1e-3, 5.0)
    opt.set_parameters([("t", t)])
    opt.step()
    t.grad = torch.zeros_like(t)
    with pytest.raises(RuntimeError):
        opt.step()


def _tem(grad_mode, seed=3):
    from prodsearch_b200 import synth
    from prodsearch_b200.item_transformer import ItemTransformerRanker
    cfg = argparse.Namespace(
        train_review_only=True, embedding_size=128, dropout=0.0, pretrain_emb_dir="", pretrain_up_emb_dir="",
        sep_prod_emb=False, model_name="item_transformer", ff_size=512, heads=8, inter_layers=1,
        query_encoder_name="fs", use_dot_prod=True, use_pos_emb=True, use_item_pos=False, sim_func="bias_product",
        pos_weight=False, neg_per_pos=5)
    torch.manual_seed(seed)
    return ItemTransformerRanker(cfg, "cuda", 6000, 9000, None, word_dists=synth.word_dists(6000), grad_mode=grad_mode)


def _cuda(ns):
    return argparse.Namespace(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(ns).items()})


@pytest.mark.parametrize("graphed", [False, True])
def test_tem_training_rowsparse_equals_dense(graphed):
    """The whole train loop -- forward (rows made current on read), backward, sinks, optimizer -- with row-sparse tables
    against the same loop with dense gradients and the dense fused Adam; then ranking (which flushes)."""
    from prodsearch_b200 import _lib, synth
    from prodsearch_b200.graph_step import GraphedTrainStep
    from prodsearch_b200.optimizers import Optimizer
    dense, sparse = _tem("dense"), _tem("rowsparse")
    sparse.load_state_dict(dense.state_dict())
    opts = []
    for m in (dense, sparse):
        o = Optimizer("adam", 5e-3, 5.0)
        o.set_parameters(list(m.named_parameters()))
        opts.append(o)
        m.train()
    P, V, B, steps = 9000, 6000, 96, 9
    batches = [synth.tem_batch(B, P, V, seed=40 + s) for s in range(steps)]
    Wq = max(b[0].query_word_idxs.shape[1] for b in batches)

    def padded(b):
        q = torch.full((B, Wq), V - 1, dtype=torch.int64)
        q[:, :b.query_word_idxs.shape[1]] = b.query_word_idxs
        return argparse.Namespace(**dict(vars(b), query_word_idxs=q))
    neg_i = torch.empty(B, 5, dtype=torch.int64, device="cuda")
    neg_w = torch.empty(B * 5, dtype=torch.int64, device="cuda")
    for m in (dense, sparse):
        m.injected_negatives = (neg_i, neg_w)
    step_fn = None
    lr = 5e-3
    for s, (b, ni, nw) in enumerate(batches):
        neg_i.copy_(ni)
        neg_w.copy_(nw)
        cb = _cuda(padded(b))
        loss = dense(cb)
        dense.zero_grad()
        loss.backward()
        opts[0].step()
        l_dense = float(loss.detach())
        del loss
        if graphed:
            if step_fn is None:
                step_fn = GraphedTrainStep(sparse, opts[1], cb)
            l_sparse = float(step_fn(cb))
        else:
            loss = sparse(cb)
            sparse.zero_grad()
            loss.backward()
            opts[1].step()
            l_sparse = float(loss.detach())
            del loss
        # the forward pass of every step reads the same (made-current) rows: same loss
        assert abs(l_sparse - l_dense) <= 2e-5 * abs(l_dense), (s, l_sparse, l_dense)
    assert sparse.product_emb.weight.grad is None and getattr(sparse.product_emb.weight, "_psb_lazy", None) is not None
    sparse.eval()                                     # flushes the resting rows
    dense.eval()
    sd_d, sd_s = dense.state_dict(), sparse.state_dict()
    # Parameters: the optimizer arithmetic itself is held to 2e-6 * steps by the kernel-level test above (gradients of
    # ordinary size).  In a real model Adam turns a gradient element that is pure rounding noise into an update of
    # +- lr whatever its size (m / sqrt(v) = +- 1 on first touch), so last-bit differences between the two flows
    # (summation order of the clip norm, series vs step-by-step decay) show up as a FEW elements that differ by a
    # fraction of lr; everything else agrees tightly.
    for k in sd_d:
        if k.endswith("linear_keys.bias"):
            continue      # its gradient is exactly 0 in exact arithmetic (softmax is shift invariant): Adam random-walks
        if sd_d[k].dtype.is_floating_point:
            err = (sd_d[k] - sd_s[k]).abs()
            assert err.max().item() <= 0.05 * lr * steps, (k, err.max().item())
            assert (err > 3e-6 * steps).float().mean().item() <= 2e-3, (k, (err > 3e-6 * steps).float().mean().item())
    cb = _cuda(padded(batches[0][0]))
    ids_d, sc_d = dense.rank_catalog(cb, k=50, mode=_lib.TOPK_EXACT)
    ids_s, sc_s = sparse.rank_catalog(cb, k=50, mode=_lib.TOPK_EXACT)
    torch.testing.assert_close(sc_d, sc_s, rtol=1e-4, atol=1e-4)
