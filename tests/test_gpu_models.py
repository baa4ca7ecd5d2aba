"""GPU parity of the drop-in modules against (a) the committed golden outputs of the reference's own
modules and (b) the CPU oracle on larger seeded batches.  Losses / scores / gradients within 1e-5
relative (north-star tolerance; gradients that are sums of many terms get 1e-4), indices exact."""
import argparse

import numpy as np
import pytest
import torch

import oracle
from golden_util import DEFAULTS, Golden

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-5, atol=2e-6):
    """Elementwise |a - b| <= atol + rtol |b| + 3e-6 max|b| (1e-5 relative in fp32 is the bar; the floor of an element
    near zero scales with the tensor it belongs to -- see tests/test_gpu_encoder.py close)."""
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    err = (a - b).abs()
    bound = atol + rtol * b.abs() + 3e-6 * (float(b.abs().max()) if b.numel() else 0.0)
    assert bool((err <= bound).all()), "max err %g at bound %g" % (
        float(err.max()), float(bound.flatten()[err.argmax()] if err.numel() else 0))


def cuda_batch(ns):
    return argparse.Namespace(**{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vars(ns).items()})


def load_params(model, G):
    sd = {}
    for k, v in G.params.items():
        if k == "review_embeddings":
            continue                                   # alias the reference registers lazily (ps_model.py:188)
        if k.endswith("pos_emb.pe"):
            v = oracle.sinusoid_table(5000, v.shape[-1])
        sd[k] = v
    model.load_state_dict(sd, strict=True)            # strict: state_dict keys are the reference's


def check_grads(model, G, rtol=1e-4, atol=2e-6):
    seen = 0
    for k, p in model.named_parameters():
        ref = G.grads.get(k)
        if ref is None:
            continue
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        close(got, ref, rtol=rtol, atol=atol)
        seen += 1
    assert seen >= 8


# ---------------------------------------------------------------- TEM vs golden
@pytest.mark.parametrize("name", ["tem_fs", "tem_avg_bias", "tem_d128", "tem_itempos"])
@pytest.mark.parametrize("grad_mode", ["dense", "rowsparse"])
def test_tem_golden(name, grad_mode):
    from prodsearch_b200.item_transformer import ItemTransformerRanker
    G = Golden(name)
    i = G.inputs
    V = G.params["word_embeddings.weight"].shape[0]
    Pn = G.params["product_emb.weight"].shape[0] - 1
    model = ItemTransformerRanker(G.cfg, "cuda", V, Pn, None, word_dists=np.ones(V), grad_mode=grad_mode)
    load_params(model, G)
    batch = cuda_batch(G.batch())
    model.train()
    model.injected_negatives = (i["neg_item_idxs"].cuda(), i["neg_word_idxs"].cuda())
    loss = model(batch)
    model.zero_grad()
    loss.backward()
    close(loss, G.outputs["loss"])
    close(model.ps_loss, G.outputs["ps_loss"])
    close(model.item_loss, G.outputs["item_loss"])
    if grad_mode == "dense":
        check_grads(model, G)
        # a second step on the same model: the persistent dense buffers must be re-zeroed correctly
        model.zero_grad()
        model(batch).backward()
        check_grads(model, G)
    else:
        for k in ("product_emb.weight", "word_embeddings.weight"):
            p = dict(model.named_parameters())[k]
            rows, vals, nu = p.row_grad
            nu = int(nu.item())
            ref = G.grads[k]
            dense = torch.zeros_like(ref)
            dense[rows[:nu].long().cpu()] = vals[:nu].cpu()
            close(dense, ref, rtol=1e-4)
            assert p.grad is None
    model.eval()
    # scores are d-term dot products of O(1) values: 1e-5 relative to the operand scale |q||e| ~ d
    close(model.test(batch), G.outputs["test_scores"], rtol=1e-5, atol=3e-5)
    # full-catalog ranking == canonical (lower-id-first) ranking of the oracle's score matrix
    k = min(10, Pn)
    ids, sc = model.rank_catalog(batch, k=k)
    P = {kk: vv for kk, vv in G.leaf_params(oracle.sinusoid_table).items()}
    with torch.no_grad():
        _, full = oracle.tem_catalog_scores(P, G.cfg, i["query_word_idxs"], i["u_item_idxs"])
    ref_i, ref_s = oracle.topk_lower_id_first(full.numpy(), k)
    close(sc, ref_s, rtol=1e-5, atol=1e-5)
    gaps = np.abs(np.diff(np.sort(full.numpy(), axis=1)[:, ::-1][:, :k + 1], axis=1)).min()
    if gaps > 1e-4:                                    # ids are only defined where score gaps exceed fp32 noise
        assert np.array_equal(ids.cpu().numpy(), ref_i)


# ---------------------------------------------------------------- TEM vs oracle at BASELINE shape
def _tem_cfg(**kw):
    c = dict(DEFAULTS)
    c.update(dict(embedding_size=128, ff_size=512, heads=8, inter_layers=1, neg_per_pos=5, model_name="item_transformer"))
    c.update(kw)
    return argparse.Namespace(**c)


@pytest.mark.parametrize("B", [384, 50])
def test_tem_vs_oracle_amazon_shape(B):
    from prodsearch_b200 import synth
    from prodsearch_b200.item_transformer import ItemTransformerRanker
    cfg = _tem_cfg()
    torch.manual_seed(5)
    P, V = 18000, 32000
    model = ItemTransformerRanker(cfg, "cuda", V, P, None, word_dists=synth.word_dists(V))
    with torch.no_grad():
        model.word_bias.normal_(0, 0.1)
    batch, neg_items, neg_words = synth.tem_batch(B, P, V, L=20, W=1, K=5, seed=17)
    params = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
    model.train()
    model.injected_negatives = (neg_items.cuda(), neg_words.cuda())
    loss = model(cuda_batch(batch))
    model.zero_grad()
    loss.backward()
    torch.set_num_threads(8)
    ref, ref_ps, ref_il = oracle.tem_forward(params, cfg, batch.query_word_idxs, batch.target_prod_idxs,
                                             batch.u_item_idxs, batch.pos_iword_idxs, neg_items, neg_words,
                                             training=True)
    ref.backward()
    close(loss, ref)
    close(model.ps_loss, ref_ps)
    close(model.item_loss, ref_il)
    for k, p in model.named_parameters():
        g = params[k].grad
        if g is None:
            continue
        g = g.clone()
        if k in ("product_emb.weight", "word_embeddings.weight"):
            g[-1] = 0
        scale = float(g.abs().max()) + 1e-12
        close(p.grad, g, rtol=1e-4, atol=max(2e-5 * scale, 2e-7))   # key-bias grads are pure fp32 noise


# ---------------------------------------------------------------- PV / PVC vs golden
def test_pv_golden(golden_dir):
    from prodsearch_b200.pv import ParagraphVector
    z = np.load(golden_dir + "/pv.npz")
    wt = torch.from_numpy(z["word_table"])
    V, d = wt.shape
    wemb = torch.nn.Embedding(V, d, padding_idx=V - 1)
    pv = ParagraphVector(wemb, torch.ones(V), z["review_table"].shape[0], dropout=0.0).cuda()
    with torch.no_grad():
        wemb.weight.copy_(wt)
        pv.review_embeddings.weight.copy_(torch.from_numpy(z["review_table"]))
    pv.injected_negatives = torch.from_numpy(z["neg_word_idxs"]).cuda()
    emb, loss = pv(torch.from_numpy(z["review_ids"]).cuda(), torch.from_numpy(z["pos_word_idxs"]).cuda(),
                   torch.from_numpy(z["word_mask"]).cuda(), int(z["n_negs"]))
    close(emb, z["review_emb"], rtol=0, atol=0)          # gathered rows: bit exact
    close(loss, z["loss"])
    ((emb * torch.from_numpy(z["up_emb"]).cuda()).sum() + (loss * torch.from_numpy(z["up_loss"]).cuda()).sum()).backward()
    close(wemb.weight.grad, z["grad_word_table"], rtol=1e-4)
    close(pv.review_embeddings.weight.grad, z["grad_review_table"], rtol=1e-4)


def test_pvc_golden(golden_dir):
    from prodsearch_b200.pvc import ParagraphVectorCorruption
    z = np.load(golden_dir + "/pvc.npz")
    wt = torch.from_numpy(z["word_table"])
    V, d = wt.shape
    wemb = torch.nn.Embedding(V, d, padding_idx=V - 1)
    pvc = ParagraphVectorCorruption(wemb, torch.ones(V), float(z["corrupt_rate"]), dropout=0.0).cuda()
    with torch.no_grad():
        wemb.weight.copy_(wt)
    pvc.injected_negatives = torch.from_numpy(z["neg_word_idxs"]).cuda()
    pvc.injected_corruption = [torch.from_numpy(z["corrupt_mask"])]
    emb, loss = pvc(torch.from_numpy(z["pos_word_idxs"]).cuda(), torch.from_numpy(z["word_mask"]).cuda(),
                    torch.from_numpy(z["rword_idxs_pvc"]).cuda(), int(z["n_negs"]))
    close(emb, z["review_emb"])
    close(loss, z["loss"])
    ((emb * torch.from_numpy(z["up_emb"]).cuda()).sum() + (loss * torch.from_numpy(z["up_loss"]).cuda()).sum()).backward()
    close(wemb.weight.grad, z["grad_word_table"], rtol=1e-4)
    pvc.injected_corruption = [torch.from_numpy(z["corrupt_mask2"])]
    with torch.no_grad():
        close(pvc.get_para_vector(torch.from_numpy(z["rword_idxs_pvc"]).cuda()), z["para_vector"])


# ---------------------------------------------------------------- RTM vs golden
@pytest.mark.parametrize("enc", ["pv", "pvc", "fs", "avg"])
@pytest.mark.parametrize("train_pv", [True, False])
def test_rtm_golden(enc, train_pv):
    from prodsearch_b200.ps_model import ProductRanker
    G = Golden("rtm_%s%s" % (enc, "_trainpv" if train_pv else ""), model_name="review_transformer",
               embedding_size=32, ff_size=48, heads=4)
    R = int(G.z["cfg/review_count"])
    V = G.params["word_embeddings.weight"].shape[0]
    Pn, U = 12, 9
    G.cfg.do_subsample_mask = True                      # review_words in the golden file are already padded
    model = ProductRanker(G.cfg, "cuda", V, R, Pn, U, G.inputs["review_words"].tolist(), None,
                          word_dists=np.ones(V))
    load_params(model, G)
    b = G.batch()
    for k in ("pos_prod_rword_idxs_pvc", "neg_prod_rword_idxs_pvc"):
        if not hasattr(b, k):
            setattr(b, k, None)
    b = cuda_batch(b)
    draws_m = [G.draws[k] for k in sorted(G.draws) if k.startswith("multinomial")]
    draws_b = [G.draws[k] for k in sorted(G.draws) if k.startswith("bernoulli")]
    if draws_m:
        model.review_encoder.injected_negatives = draws_m[0].cuda()
    if draws_b:
        model.review_encoder.injected_corruption = list(draws_b)
    model.train()
    loss = model(b, train_pv=train_pv)
    model.zero_grad()
    loss.backward()
    close(loss, G.outputs["loss"])
    for alias, real in (("review_embeddings", "review_encoder.review_embeddings.weight"),):
        if alias in G.grads:
            G.grads.setdefault(real, G.grads.pop(alias))
    check_grads(model, G)
    model.eval()
    model.get_review_embeddings()
    close(model.review_embeddings, G.outputs["review_table"])
    close(model.test(b), G.outputs["test_scores"])
    model.clear_review_embbeddings()
