"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors, so the
pin for the oracle is the output of the reference's own modules on small seeded
inputs: the modules are imported unmodified from /root/reference, with
  * a uint8->bool cast shim around Tensor.masked_fill/masked_fill_ (the reference
    predates torch 2.x bool masks), and
  * torch.multinomial / torch.bernoulli replaced by queues of pre-drawn tensors so
    the sampled negatives and the PVC corruption mask are inputs, not RNG state.
Each .npz holds inputs (parameters, batch, supplied draws) and outputs (losses,
scores, dense gradients).  Nothing here is imported by the product.
"""
import argparse
import contextlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- shims
def install_mask_shim():
    for name in ("masked_fill", "masked_fill_"):
        orig = getattr(torch.Tensor, name)

        def patched(self, mask, value, _orig=orig):
            if isinstance(mask, torch.Tensor) and mask.dtype == torch.uint8:
                mask = mask.bool()
            return _orig(self, mask, value)
        setattr(torch.Tensor, name, patched)


@contextlib.contextmanager
def injected(multinomial=(), bernoulli=()):
    """Replace the RNG draws inside forward() by supplied tensors, in call order."""
    mq, bq = list(multinomial), list(bernoulli)
    o_m, o_b = torch.multinomial, torch.bernoulli

    def fake_m(dist, n, replacement=False):
        out = mq.pop(0)
        assert out.numel() == n, (out.numel(), n)
        return out.reshape(-1).clone()

    def fake_b(probs, *a, **k):
        out = bq.pop(0)
        assert out.shape == probs.shape, (out.shape, probs.shape)
        return out.to(probs.dtype).clone()
    torch.multinomial, torch.bernoulli = fake_m, fake_b
    try:
        yield
    finally:
        torch.multinomial, torch.bernoulli = o_m, o_b
    assert not mq and not bq, "unused injected draws"


def base_args(**kw):
    a = argparse.Namespace(
        train_review_only=True, embedding_size=64, dropout=0.0, pretrain_emb_dir="",
        pretrain_up_emb_dir="", sep_prod_emb=False, model_name="item_transformer", ff_size=64,
        heads=8, inter_layers=1, query_encoder_name="fs", use_dot_prod=True, use_pos_emb=True,
        use_item_pos=False, sim_func="product", pos_weight=False, neg_per_pos=3,
        review_encoder_name="pv", fix_emb=False, do_subsample_mask=False, review_word_limit=6,
        use_user_emb=False, use_item_emb=False, use_seg_emb=True, corrupt_rate=0.5)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def np_state(model):
    """All parameters; the 5000-row sinusoid buffer is cut to its first 64 rows (the
    oracle regenerates it with sinusoid_table and the test checks the slice)."""
    out = {}
    for k, v in model.state_dict().items():
        v = v.detach().numpy().copy()
        out["param/" + k] = v[:, :64] if k.endswith("pos_emb.pe") else v
    return out


def np_grads(model):
    return {"grad/" + k: (p.grad.detach().numpy().copy() if p.grad is not None
                          else np.zeros(tuple(p.shape), np.float32))
            for k, p in model.named_parameters()}


class Obj(object):
    pass


# ---------------------------------------------------------------- TEM
def make_tem(name, seed, **kw):
    from models.item_transformer import ItemTransformerRanker
    from data.batch_data import ItemPVBatch
    g = torch.Generator().manual_seed(seed)
    args = base_args(**kw)
    V, Pn, B, L, Wq, W, K, C = 40, 30, 6, 5, 4, 2, args.neg_per_pos, 12
    word_dists = (np.arange(1, V + 1, dtype=np.float64) ** -0.75)
    word_dists[-1] = 0
    word_dists = (word_dists / word_dists.sum()).tolist()
    torch.manual_seed(seed)
    model = ItemTransformerRanker(args, "cpu", V, Pn, None, word_dists=word_dists)
    with torch.no_grad():   # biases start at zero in the reference; make them matter
        model.product_bias.copy_(torch.randn(Pn + 1, generator=g) * 0.3)
        model.word_bias.copy_(torch.randn(V, generator=g) * 0.3)
        model.query_encoder.f_W.bias.copy_(torch.randn(args.embedding_size, generator=g) * 0.1) \
            if args.query_encoder_name == "fs" else None
    qw = torch.randint(0, V - 1, (B, Wq), generator=g)
    qw[0, 2:] = V - 1
    qw[3, 1:] = V - 1
    tgt = torch.randint(0, Pn, (B,), generator=g)
    hist = torch.randint(0, Pn, (B, L), generator=g)
    hist[1, 3:] = Pn
    hist[4, :] = Pn                      # a user with no history at all
    hist[2, 4] = hist[2, 0]              # duplicate rows inside one history
    iw = torch.randint(0, V - 1, (B, W), generator=g)
    iw[2, 1] = V - 1                     # padded target word
    iw[5, :] = V - 1                     # all target words padded -> count clamps to 1
    neg_items = torch.randint(0, Pn, (B, K), generator=g)
    neg_items[0, 0] = tgt[0]             # negative collides with the positive
    neg_words = torch.randint(0, V - 1, (B * W * K,), generator=g)
    cand = torch.randint(0, Pn, (B, C), generator=g)
    cand[:, -2:] = Pn                    # padded candidates (score 0 (+bias[pad]))
    batch = ItemPVBatch(qw, tgt, hist, iw, candi_prod_idxs=cand, to_tensor=False)
    model.train()
    with injected(multinomial=[neg_items, neg_words]):
        loss = model(batch)
    model.zero_grad()
    loss.backward()
    ps_loss, item_loss = model.ps_loss, model.item_loss
    model.eval()
    with torch.no_grad():
        test_scores = model.test(batch)
    out = dict(np_state(model))
    out.update(np_grads(model))
    out.update({
        "cfg/keys": np.array(sorted(kw.keys())), "cfg/vals": np.array([str(kw[k]) for k in sorted(kw.keys())]),
        "in/query_word_idxs": qw.numpy(), "in/target_prod_idxs": tgt.numpy(), "in/u_item_idxs": hist.numpy(),
        "in/pos_iword_idxs": iw.numpy(), "in/neg_item_idxs": neg_items.numpy(),
        "in/neg_word_idxs": neg_words.numpy(), "in/candi_prod_idxs": cand.numpy(),
        "out/loss": loss.detach().numpy(), "out/ps_loss": np.float32(ps_loss),
        "out/item_loss": np.float32(item_loss), "out/test_scores": test_scores.numpy(),
        "out/reference_rank": test_scores.numpy().argsort(axis=-1)[:, ::-1].copy(),
    })
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, float(loss), ps_loss, item_loss)


# ---------------------------------------------------------------- encoders / PV / PVC
def make_text(seed=11):
    from models.text_encoder import get_vector_mean, FSEncoder, AVGEncoder
    torch.manual_seed(seed)            # parameters come from the global RNG: seeded, so the file reproduces bit for bit
    g = torch.Generator().manual_seed(seed)
    N, W, d = 7, 5, 128
    x = torch.randn(N, W, d, generator=g, requires_grad=True)
    mask = torch.rand(N, W, generator=g) > 0.4
    mask[2] = False
    mean = get_vector_mean(x, mask)
    fs = FSEncoder(d, 0.0)
    fs.initialize_parameters()
    with torch.no_grad():
        fs.f_W.bias.copy_(torch.randn(d, generator=g) * 0.1)
    y = fs(x, mask)
    up = torch.randn(N, d, generator=g)
    (y * up).sum().backward()
    avg = AVGEncoder(d, 0.0)(x, mask)
    np.savez_compressed(os.path.join(OUT, "text_encoder.npz"), x=x.detach().numpy(), mask=mask.numpy(),
                        mean=mean.detach().numpy(), fs_w=fs.f_W.weight.detach().numpy(),
                        fs_b=fs.f_W.bias.detach().numpy(), fs_out=y.detach().numpy(), upstream=up.numpy(),
                        grad_x=x.grad.numpy(), grad_w=fs.f_W.weight.grad.numpy(),
                        grad_b=fs.f_W.bias.grad.numpy(), avg_out=avg.detach().numpy())
    print("text_encoder ok")


def make_pv(seed=21):
    from models.PV import ParagraphVector
    torch.manual_seed(seed)            # parameters come from the global RNG: seeded, so the file reproduces bit for bit
    g = torch.Generator().manual_seed(seed)
    V, R, N, W, K, d = 50, 41, 10, 3, 4, 128
    wemb = torch.nn.Embedding(V, d, padding_idx=V - 1)
    torch.nn.init.normal_(wemb.weight)
    pv = ParagraphVector(wemb, torch.ones(V), R, dropout=0.0)
    pv.initialize_parameters()
    rid = torch.randint(0, R - 1, (N,), generator=g)
    rid[3] = R - 1
    rid[5] = rid[0]
    pw = torch.randint(0, V - 1, (N, W), generator=g)
    pw[1, 2] = V - 1
    pw[3, :] = V - 1
    wmask = pw.ne(V - 1).byte()
    negs = torch.randint(0, V - 1, (N * W * K,), generator=g)
    with injected(multinomial=[negs]):
        emb, loss = pv(rid, wemb(pw), wmask, K)
    up_e = torch.randn(N, d, generator=g)
    up_l = torch.randn(N, 1, generator=g)
    ((emb * up_e).sum() + (loss * up_l).sum()).backward()
    np.savez_compressed(os.path.join(OUT, "pv.npz"), word_table=wemb.weight.detach().numpy(),
                        review_table=pv.review_embeddings.weight.detach().numpy(), review_ids=rid.numpy(),
                        pos_word_idxs=pw.numpy(), word_mask=wmask.numpy(), neg_word_idxs=negs.numpy(),
                        n_negs=K, review_emb=emb.detach().numpy(), loss=loss.detach().numpy(),
                        up_emb=up_e.numpy(), up_loss=up_l.numpy(), grad_word_table=wemb.weight.grad.numpy(),
                        grad_review_table=pv.review_embeddings.weight.grad.numpy())
    print("pv ok", float(loss.sum()))


def make_pvc(seed=31):
    from models.PVC import ParagraphVectorCorruption
    torch.manual_seed(seed)            # parameters come from the global RNG: seeded, so the file reproduces bit for bit
    g = torch.Generator().manual_seed(seed)
    V, N, W, K, Wr, d, rate = 50, 9, 2, 3, 12, 128, 0.5
    wemb = torch.nn.Embedding(V, d, padding_idx=V - 1)
    torch.nn.init.normal_(wemb.weight)
    pvc = ParagraphVectorCorruption(wemb, torch.ones(V), rate, dropout=0.0)
    rw = torch.randint(0, V - 1, (N, Wr), generator=g)
    rw[0, 7:] = V - 1
    rw[4, :] = V - 1
    pw = torch.randint(0, V - 1, (N, W), generator=g)
    pw[2, 1] = V - 1
    wmask = pw.ne(V - 1).byte()
    negs = torch.randint(0, V - 1, (N * W * K,), generator=g)
    cmask = (torch.rand(N, Wr, generator=g) < rate).float()
    with injected(multinomial=[negs], bernoulli=[cmask]):
        emb, loss = pvc(wemb(pw), wmask, rw, K)
    up_e = torch.randn(N, d, generator=g)
    up_l = torch.randn(N, 1, generator=g)
    ((emb * up_e).sum() + (loss * up_l).sum()).backward()
    cmask2 = (torch.rand(N, Wr, generator=g) < rate).float()
    with injected(bernoulli=[cmask2]), torch.no_grad():
        para = pvc.get_para_vector(rw)
    np.savez_compressed(os.path.join(OUT, "pvc.npz"), word_table=wemb.weight.detach().numpy(),
                        rword_idxs_pvc=rw.numpy(), pos_word_idxs=pw.numpy(), word_mask=wmask.numpy(),
                        neg_word_idxs=negs.numpy(), n_negs=K, corrupt_rate=rate, corrupt_mask=cmask.numpy(),
                        review_emb=emb.detach().numpy(), loss=loss.detach().numpy(), up_emb=up_e.numpy(),
                        up_loss=up_l.numpy(), grad_word_table=wemb.weight.grad.numpy(),
                        corrupt_mask2=cmask2.numpy(), para_vector=para.numpy())
    print("pvc ok", float(loss.sum()))


# ---------------------------------------------------------------- RTM
def make_rtm(name, seed, train_pv, **kw):
    from models.ps_model import ProductRanker
    from data.batch_data import ProdSearchTrainBatch, ProdSearchTestBatch
    g = torch.Generator().manual_seed(seed)
    args = base_args(model_name="review_transformer", embedding_size=32, ff_size=48, heads=4, **kw)
    V, R, Pn, U, B, K = 40, 25, 12, 9, 4, args.neg_per_pos
    Ru, Ri, Wr, Wq, C = 2, 3, args.review_word_limit, 4, 5
    Rc = Ru + Ri
    enc = args.review_encoder_name
    W = 2 if (train_pv and "pv" in enc) else Wr
    review_words = [[int(x) for x in torch.randint(0, V - 1, (int(n),), generator=g)]
                    for n in torch.randint(1, Wr + 3, (R - 1,), generator=g)]
    review_words.append([V - 1] * Wr)
    word_dists = torch.ones(V).tolist()
    torch.manual_seed(seed)
    model = ProductRanker(args, "cpu", V, R, Pn, U, review_words, None, word_dists=word_dists)

    def ridx(*shape):
        x = torch.randint(0, R - 1, shape, generator=g)
        return x
    qw = torch.randint(0, V - 1, (B, Wq), generator=g)
    qw[1, 2:] = V - 1
    pos_r = ridx(B, Rc)
    pos_r[0, 4:] = R - 1
    pos_r[2, 1] = R - 1
    neg_r = ridx(B, K, Rc)
    neg_r[1, 0, 3:] = R - 1
    neg_r[3, 2, Ru:] = R - 1
    neg_r[3, 2, :Ru] = R - 1            # a negative whose reviews are all padding -> weight 0
    seg = torch.tensor([0] + [1] * Ru + [2] * Ri)
    pos_seg = seg.unsqueeze(0).expand(B, -1).clone()
    neg_seg = seg.view(1, 1, -1).expand(B, K, -1).clone()
    pos_seg[:, 1:][pos_r == R - 1] = 3
    neg_seg[:, :, 1:][neg_r == R - 1] = 3
    table = torch.tensor(model.review_words.tolist())
    pos_rw = torch.randint(0, V - 1, (B, Rc, W), generator=g) if W != Wr else table[pos_r]
    if W != Wr:
        pos_rw[pos_r == R - 1] = V - 1
        pos_rw[1, 1, 1] = V - 1
    pos_rw_mask = pos_rw.ne(V - 1).byte()
    neg_rw = table[neg_r]
    neg_rw_mask = neg_rw.ne(V - 1).byte()
    pos_u = torch.randint(0, U, (B, Rc + 1), generator=g)
    neg_u = torch.randint(0, U, (B, K, Rc + 1), generator=g)
    pos_i = torch.randint(0, Pn, (B, Rc + 1), generator=g)
    neg_i = torch.randint(0, Pn, (B, K, Rc + 1), generator=g)
    pos_pvc = table[pos_r] if enc == "pvc" else None
    neg_pvc = table[neg_r] if enc == "pvc" else None
    batch = ProdSearchTrainBatch(qw, pos_r, pos_seg, pos_rw, pos_rw_mask, neg_r, neg_seg, pos_u, neg_u,
                                 pos_i, neg_i, neg_rw, neg_rw_mask, pos_pvc, neg_pvc, to_tensor=False)
    draws_m, draws_b = [], []
    rate = args.corrupt_rate
    if "pv" in enc and train_pv:
        draws_m.append(torch.randint(0, V - 1, (B * Rc * W * K,), generator=g))
    if enc == "pvc":
        draws_b.append((torch.rand(B * Rc, pos_pvc.size(-1) if train_pv else W, generator=g) < rate).float())
        draws_b.append((torch.rand(B * K * Rc, Wr, generator=g) < rate).float())
    model.train()
    with injected(multinomial=draws_m, bernoulli=draws_b):
        loss = model(batch, train_pv=train_pv)
    model.zero_grad()
    loss.backward()
    # test path
    model.eval()
    cand_r = ridx(B, C, Rc)
    cand_r[0, 1, 2:] = R - 1
    cand_seg = seg.view(1, 1, -1).expand(B, C, -1).clone()
    cand_seg[:, :, 1:][cand_r == R - 1] = 3
    cand_u = torch.randint(0, U, (B, C, Rc + 1), generator=g)
    cand_i = torch.randint(0, Pn, (B, C, Rc + 1), generator=g)
    tb = ProdSearchTestBatch([0] * B, [0] * B, [0] * B, [[0] * C] * B, qw, cand_r, cand_seg, cand_u, cand_i,
                             to_tensor=False)
    with torch.no_grad():
        model.get_review_embeddings()
        review_table = model.review_embeddings.detach().clone()
        test_scores = model.test(tb)
    out = dict(np_state(model))
    out.update(np_grads(model))
    out.update({
        "cfg/keys": np.array(sorted(kw.keys())), "cfg/vals": np.array([str(kw[k]) for k in sorted(kw.keys())]),
        "cfg/train_pv": np.array(train_pv), "cfg/review_count": np.array(R),
        "in/review_words": model.review_words.numpy(),
        "in/query_word_idxs": qw.numpy(), "in/pos_prod_ridxs": pos_r.numpy(), "in/pos_seg_idxs": pos_seg.numpy(),
        "in/pos_prod_rword_idxs": pos_rw.numpy(), "in/pos_prod_rword_masks": pos_rw_mask.numpy(),
        "in/neg_prod_ridxs": neg_r.numpy(), "in/neg_seg_idxs": neg_seg.numpy(),
        "in/pos_user_idxs": pos_u.numpy(), "in/neg_user_idxs": neg_u.numpy(),
        "in/pos_item_idxs": pos_i.numpy(), "in/neg_item_idxs": neg_i.numpy(),
        "in/neg_prod_rword_idxs": neg_rw.numpy(), "in/neg_prod_rword_masks": neg_rw_mask.numpy(),
        "in/candi_prod_ridxs": cand_r.numpy(), "in/candi_seg_idxs": cand_seg.numpy(),
        "in/candi_seq_user_idxs": cand_u.numpy(), "in/candi_seq_item_idxs": cand_i.numpy(),
        "out/loss": loss.detach().numpy(), "out/review_table": review_table.numpy(),
        "out/test_scores": test_scores.numpy(),
    })
    if pos_pvc is not None:
        out["in/pos_prod_rword_idxs_pvc"] = pos_pvc.numpy()
        out["in/neg_prod_rword_idxs_pvc"] = neg_pvc.numpy()
    for i, t in enumerate(draws_m):
        out["draw/multinomial%d" % i] = t.numpy()
    for i, t in enumerate(draws_b):
        out["draw/bernoulli%d" % i] = t.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, float(loss))


def make_rank(seed=41):
    """The reference's literal ranking expression on tie-free and tied scores
    (trainer.py:152); documents that ties are NOT lower-id-first there."""
    g = np.random.default_rng(seed)
    s = g.standard_normal((5, 300)).astype(np.float32)
    tied = np.array([[1, 3, 3, 2, 3, 0]], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "rank.npz"), scores=s, order=s.argsort(axis=-1)[:, ::-1].copy(),
                        tied=tied, tied_order=tied.argsort(axis=-1)[:, ::-1].copy())
    print("rank ok")


if __name__ == "__main__":
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference exists"
    sys.path.insert(0, REF)
    torch.set_num_threads(1)
    install_mask_shim()
    make_text()
    make_pv()
    make_pvc()
    make_tem("tem_fs", 101)
    make_tem("tem_avg_bias", 102, query_encoder_name="avg", sim_func="bias_product", pos_weight=True,
             sep_prod_emb=True, inter_layers=2, embedding_size=32, heads=4, ff_size=48)
    make_tem("tem_d128", 103, embedding_size=128, ff_size=32, neg_per_pos=5)
    make_tem("tem_itempos", 104, use_item_pos=True, use_pos_emb=False, neg_per_pos=4)   # last (maybe padded) position
    for enc in ("pv", "pvc", "fs", "avg"):
        make_rtm("rtm_%s_trainpv" % enc, 200 + len(enc), True, review_encoder_name=enc)
        make_rtm("rtm_%s" % enc, 300 + len(enc), False, review_encoder_name=enc,
                 use_user_emb=(enc == "pv"), use_item_emb=(enc == "pv"), query_encoder_name="avg")
    make_rank()
