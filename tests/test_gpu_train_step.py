"""GPU: the fused clipped Adam (psb_adam_step) against torch.optim.Adam(eps=1e-9) + clip_grad_norm_ -- the
reference's optimizer step (models/optimizers.py:205-243) -- and the CUDA-graph train step against the same
steps launched eagerly (bit-identical parameters: the graph replays exactly the eager kernel sequence)."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk_params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1001, 128), (128, 128), (128,), (513,), (37, 5), (3,), (70000,)]
    return [torch.randn(s, generator=g) for s in shapes]


@pytest.mark.parametrize("max_norm,noam,wd", [(5.0, False, 0.0), (0.05, False, 0.0), (0.0, False, 0.0),
                                              (5.0, True, 0.0), (1.0, False, 0.01)])
def test_fused_adam_matches_torch(max_norm, noam, wd):
    from prodsearch_b200.optimizers import Optimizer
    init = _mk_params(1)
    ref_p = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    my_p = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    lr, warm = 5e-4, 20
    ref_opt = torch.optim.Adam(ref_p, lr=lr, betas=(0.9, 0.999), eps=1e-9, weight_decay=wd)
    opt = Optimizer("adam", lr, max_norm, decay_method="noam" if noam else "adam", warmup_steps=warm, weight_decay=wd)
    opt.set_parameters([("p%d" % i, p) for i, p in enumerate(my_p)])
    g = torch.Generator().manual_seed(2)
    for step in range(1, 8):
        grads = [torch.randn(t.shape, generator=g) * (0.01 if step % 2 else 3.0) for t in init]
        if step == 3:
            grads[0][5:] = 0          # mostly-zero embedding gradient
        for p, q, gr in zip(ref_p, my_p, grads):
            p.grad = gr.clone().cuda()
            q.grad = gr.clone().cuda()
        if noam:
            ref_opt.param_groups[0]["lr"] = lr * min(step ** -0.5, step * warm ** -1.5)
        if max_norm:
            total = torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        ref_opt.step()
        opt.step()
        if max_norm:
            assert abs(float(opt.optimizer.total_norm) - float(total)) <= 1e-5 * float(total)
        for p, q in zip(ref_p, my_p):
            err = (p.detach() - q.detach()).abs().max().item()
            assert err <= 2e-6 * step, (step, err)
    sd = opt.optimizer.state_dict()
    assert set(sd) == {"state", "param_groups"} and float(sd["state"][0]["step"]) == 7
    ref_sd = ref_opt.state_dict()
    for i in range(len(init)):
        assert torch.allclose(sd["state"][i]["exp_avg"], ref_sd["state"][i]["exp_avg"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"], rtol=1e-5, atol=1e-7)
    # round trip through the torch-format checkpoint
    opt2 = Optimizer("adam", lr, max_norm)
    opt2.set_parameters([("p%d" % i, p) for i, p in enumerate(my_p)])
    opt2.optimizer.load_state_dict(sd)
    assert int(opt2.optimizer._step_dev.item()) == 7


def _tem(dropout, seed=0):
    from prodsearch_b200 import synth
    from prodsearch_b200.item_transformer import ItemTransformerRanker
    from prodsearch_b200.optimizers import build_optim
    cfg = argparse.Namespace(
        train_review_only=True, embedding_size=128, dropout=dropout, pretrain_emb_dir="", pretrain_up_emb_dir="",
        sep_prod_emb=False, model_name="item_transformer", ff_size=512, heads=8, inter_layers=1,
        query_encoder_name="fs", use_dot_prod=True, use_pos_emb=True, use_item_pos=False, sim_func="product",
        pos_weight=False, neg_per_pos=5, optim="adam", lr=0.0005, max_grad_norm=5.0, beta1=0.9, beta2=0.999,
        decay_method="adam", warmup_steps=8000, l2_lambda=0.0, train_from="")
    torch.manual_seed(seed)
    P, V = 3000, 5000
    model = ItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V))
    return model, build_optim(cfg, model), cfg, P, V


def test_graph_step_equals_eager_steps():
    from prodsearch_b200 import synth
    from prodsearch_b200.graph_step import GraphedTrainStep
    B = 96
    batches = []
    for it in range(4):
        b, ni, nw = synth.tem_batch(B, 3000, 5000, seed=50 + it, Wq_max=9)
        batches.append((b, ni.cuda(), nw.cuda()))
    widest = max(b.query_word_idxs.shape[1] for b, _, _ in batches)

    def padded(b):   # the eager run sees the same right-padded query matrix the graph's static buffer holds
        q = torch.full((B, widest), 5000 - 1, dtype=torch.int64)
        q[:, :b.query_word_idxs.shape[1]] = b.query_word_idxs
        out = argparse.Namespace(**vars(b))
        out.query_word_idxs = q
        return out

    def run(graphed, packed=False):
        model, optim, cfg, P, V = _tem(0.0, seed=7)
        model.train()
        neg_i = batches[0][1].clone()
        neg_w = batches[0][2].clone()
        model.injected_negatives = (neg_i, neg_w)
        losses = []
        if graphed:
            sample = argparse.Namespace(**vars(batches[0][0]))
            sample.query_word_idxs = torch.full((B, widest), V - 1, dtype=torch.int64)
            step = GraphedTrainStep(model, optim, sample, pad_values={"query_word_idxs": V - 1, "u_item_idxs": P})
            assert step.launches_per_replay >= 10
        for b, ni, nw in batches:
            neg_i.copy_(ni)
            neg_w.copy_(nw)
            if graphed:      # plain batch: one copy per field; packed: the narrower batch is padded on the host, one copy
                losses.append(float(step(step.pack(b) if packed else b)))
            else:
                db = argparse.Namespace(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(padded(b)).items()})
                loss = model(db)
                model.zero_grad()
                loss.backward()
                optim.step()
                losses.append(float(loss))
        return losses, [p.detach().clone() for p in model.parameters()], model.ps_loss

    l_e, p_e, ps_e = run(False)
    l_g, p_g, ps_g = run(True)
    assert l_e == l_g, (l_e, l_g)
    assert abs(ps_e - ps_g) <= 1e-6 * abs(ps_e)
    for a, b in zip(p_e, p_g):
        assert torch.equal(a, b)
    l_p, p_p, _ = run(True, packed=True)
    assert l_p == l_g
    for a, b in zip(p_p, p_g):
        assert torch.equal(a, b)


def test_graph_step_with_dropout_and_sampled_negatives_trains():
    """Replays draw fresh negatives / dropout masks from the device generator: losses differ step to step
    on the SAME batch, and the loss goes down."""
    from prodsearch_b200 import synth
    from prodsearch_b200.graph_step import GraphedTrainStep
    model, optim, cfg, P, V = _tem(0.1, seed=3)
    model.train()
    b, _, _ = synth.tem_batch(128, P, V, seed=9)
    step = GraphedTrainStep(model, optim, b)
    losses = [float(step()) for _ in range(60)]
    assert len(set(losses[:5])) == 5
    assert np.mean(losses[-10:]) < np.mean(losses[:10])
    assert all(np.isfinite(losses))
