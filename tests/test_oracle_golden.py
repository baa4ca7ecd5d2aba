"""CPU: the oracle restatement against the reference's own outputs (tests/golden).

This is the pin SURVEY.md 8(c) asks for: the reference has no tests of its own, so
the golden files hold what its modules produced in the build container."""
import numpy as np
import pytest
import torch

import oracle
from golden_util import Golden

torch.set_num_threads(1)
TOL = dict(rtol=2e-5, atol=2e-6)


def close(a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    assert torch.allclose(torch.as_tensor(a), torch.as_tensor(b), **tol), \
        float((torch.as_tensor(a) - torch.as_tensor(b)).abs().max())


def test_text_encoder():
    z = np.load(Golden.__init__.__globals__["GOLDEN"] + "/text_encoder.npz")
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    mask = torch.from_numpy(z["mask"])
    close(oracle.masked_mean(x, mask), z["mean"])
    close(oracle.avg_encoder(x, mask), z["avg_out"])
    w = torch.from_numpy(z["fs_w"]).requires_grad_(True)
    b = torch.from_numpy(z["fs_b"]).requires_grad_(True)
    y = oracle.fs_encoder(x, mask, w, b)
    close(y, z["fs_out"])
    (y * torch.from_numpy(z["upstream"])).sum().backward()
    close(x.grad, z["grad_x"])
    close(w.grad, z["grad_w"])
    close(b.grad, z["grad_b"])


def test_pv():
    z = np.load(Golden.__init__.__globals__["GOLDEN"] + "/pv.npz")
    wt = torch.from_numpy(z["word_table"]).requires_grad_(True)
    rt = torch.from_numpy(z["review_table"]).requires_grad_(True)
    emb, loss = oracle.pv_forward(rt, wt, torch.from_numpy(z["review_ids"]), torch.from_numpy(z["pos_word_idxs"]),
                                  torch.from_numpy(z["word_mask"]), torch.from_numpy(z["neg_word_idxs"]),
                                  int(z["n_negs"]))
    close(emb, z["review_emb"])
    close(loss, z["loss"])
    ((emb * torch.from_numpy(z["up_emb"])).sum() + (loss * torch.from_numpy(z["up_loss"])).sum()).backward()
    gw, gr = wt.grad.clone(), rt.grad.clone()
    gw[-1] = 0   # nn.Embedding(padding_idx=...) zeroes the pad-row gradient (PV.py:34-35)
    gr[-1] = 0
    close(gw, z["grad_word_table"])
    close(gr, z["grad_review_table"])


def test_pvc():
    z = np.load(Golden.__init__.__globals__["GOLDEN"] + "/pvc.npz")
    wt = torch.from_numpy(z["word_table"]).requires_grad_(True)
    emb, loss = oracle.pvc_forward(wt, wt, torch.from_numpy(z["pos_word_idxs"]), torch.from_numpy(z["word_mask"]),
                                   torch.from_numpy(z["rword_idxs_pvc"]), torch.from_numpy(z["neg_word_idxs"]),
                                   int(z["n_negs"]), torch.from_numpy(z["corrupt_mask"]), float(z["corrupt_rate"]))
    close(emb, z["review_emb"])
    close(loss, z["loss"])
    ((emb * torch.from_numpy(z["up_emb"])).sum() + (loss * torch.from_numpy(z["up_loss"])).sum()).backward()
    gw = wt.grad.clone()
    gw[-1] = 0
    close(gw, z["grad_word_table"])
    para = oracle.pvc_para_vector(wt.detach(), torch.from_numpy(z["rword_idxs_pvc"]), wt.shape[0] - 1,
                                  torch.from_numpy(z["corrupt_mask2"]), float(z["corrupt_rate"]))
    close(para, z["para_vector"])


def _zero_pad_rows(P, grads):
    """padding_idx rows get a zero gradient in the reference (nn.Embedding semantics)."""
    out = {}
    for k, p in P.items():
        if not (torch.is_tensor(p) and p.requires_grad):
            continue
        g = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
        if k in ("product_emb.weight", "hist_product_emb.weight", "word_embeddings.weight",
                 "seg_embeddings.weight", "user_emb.weight", "review_encoder.review_embeddings.weight"):
            g[-1] = 0
        out[k] = g
    return out


@pytest.mark.parametrize("name", ["tem_fs", "tem_avg_bias", "tem_d128", "tem_itempos"])
def test_tem(name):
    G = Golden(name)
    P = G.leaf_params(oracle.sinusoid_table)
    i = G.inputs
    loss, ps, il = oracle.tem_forward(P, G.cfg, i["query_word_idxs"], i["target_prod_idxs"], i["u_item_idxs"],
                                      i["pos_iword_idxs"], i["neg_item_idxs"], i["neg_word_idxs"], training=True)
    close(loss, G.outputs["loss"])
    close(ps, G.outputs["ps_loss"])
    close(il, G.outputs["item_loss"])
    loss.backward()
    for k, g in _zero_pad_rows(P, G.grads).items():
        close(g, G.grads[k], rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        sc = oracle.tem_test_scores(P, G.cfg, i["query_word_idxs"], i["u_item_idxs"], i["candi_prod_idxs"])
        close(sc, G.outputs["test_scores"])
        # the single-GEMM restatement agrees with the per-candidate re-encode
        q, full = oracle.tem_catalog_scores(P, G.cfg, i["query_word_idxs"], i["u_item_idxs"],
                                            n_items=P["product_emb.weight"].shape[0])
        close(torch.gather(full, 1, i["candi_prod_idxs"]), G.outputs["test_scores"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("enc", ["pv", "pvc", "fs", "avg"])
@pytest.mark.parametrize("train_pv", [True, False])
def test_rtm(enc, train_pv):
    G = Golden("rtm_%s%s" % (enc, "_trainpv" if train_pv else ""), model_name="review_transformer",
               embedding_size=32, ff_size=48, heads=4)
    G.cfg.review_pad_idx = int(G.z["cfg/review_count"]) - 1
    # the reference registers aliases of shared tensors (ps_model.py:188, PV.py:19); keep one name
    for alias, real in (("review_embeddings", "review_encoder.review_embeddings.weight"),
                        ("review_encoder.word_embeddings.weight", "word_embeddings.weight"),
                        ("review_encoder.context_embeddings.weight", "word_embeddings.weight")):
        G.params.pop(alias, None)
        if alias in G.grads:
            G.grads.setdefault(real, G.grads[alias])
            G.grads.pop(alias)
    P = G.leaf_params(oracle.sinusoid_table)
    b = G.batch()
    for k in ("pos_prod_rword_idxs_pvc", "neg_prod_rword_idxs_pvc"):
        if not hasattr(b, k):
            setattr(b, k, None)
    draws_m = [G.draws[k] for k in sorted(G.draws) if k.startswith("multinomial")]
    draws_b = [G.draws[k] for k in sorted(G.draws) if k.startswith("bernoulli")]
    loss, _, _ = oracle.rtm_forward(P, G.cfg, b, train_pv, draws_m[0] if draws_m else None, draws_b, training=True)
    close(loss, G.outputs["loss"])
    loss.backward()
    for k, g in _zero_pad_rows(P, G.grads).items():
        close(g, G.grads[k], rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        table = oracle.rtm_review_embeddings(P, G.cfg, G.inputs["review_words"])
        close(table, G.outputs["review_table"])
        close(oracle.rtm_test_scores(P, G.cfg, table, b), G.outputs["test_scores"])


def test_rank_contract():
    z = np.load(Golden.__init__.__globals__["GOLDEN"] + "/rank.npz")
    # tie-free: lower-id-first canonical order == the reference's literal expression
    assert np.array_equal(oracle.rank_lower_id_first(z["scores"]), z["order"])
    assert np.array_equal(oracle.reference_rank(z["scores"]), z["order"])
    ids, sc = oracle.topk_lower_id_first(z["scores"], 100)
    assert np.array_equal(ids, z["order"][:, :100])
    # tied: the reference is NOT lower-id-first; the contract is
    assert z["tied_order"].tolist() == [[4, 2, 1, 3, 0, 5]]
    assert oracle.rank_lower_id_first(z["tied"]).tolist() == [[1, 2, 4, 3, 0, 5]]
    # sharded merge == global
    s = z["scores"]
    parts = [s[:, g::3] for g in range(3)]
    pid = [np.tile(np.arange(g, s.shape[1], 3), (s.shape[0], 1)) for g in range(3)]
    loc = [oracle.topk_lower_id_first(p, 50, i) for p, i in zip(parts, pid)]
    mi, ms = oracle.merge_shard_topk([l[0] for l in loc], [l[1] for l in loc], 50)
    assert np.array_equal(mi, z["order"][:, :50])
