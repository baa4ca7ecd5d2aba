"""GPU: kernel-level checks of the entry points added in round 2, each against the older entry point (or a plain torch
restatement) it must agree with bit for bit / within fp32 rounding:

  psb_tem_loss_fwd / psb_tem_loss_finish   == psb_ns_loss_fwd on the repacked rows + mean / sum
  psb_scatter_sort_rows + psb_scatter_reduce_sorted  == psb_scatter_reduce_rows (bitwise, all sort paths)
  psb_grad_sqnorm_sparse (compact and by-row gradients)  == sum of squares in fp64
  psb_peer_gather_rows_lazy  == psb_adam_rows_catchup on the owner, then psb_peer_gather_rows
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from prodsearch_b200 import ops as o
    return o


@pytest.mark.parametrize("B,K,d,with_bias,pos_weight", [(384, 5, 128, False, 1.0), (77, 7, 128, True, 5.0),
                                                        (50, 3, 64, True, 1.0), (1, 1, 32, False, 1.0)])
def test_tem_loss_on_the_encoder_block_equals_ns_loss_on_repacked_rows(ops, B, K, d, with_bias, pos_weight):
    g = torch.Generator(device="cuda").manual_seed(B + K)
    rows = 5000
    table = torch.randn(rows, d, device="cuda", generator=g)
    bias = torch.randn(rows, device="cuda", generator=g) * 0.1 if with_bias else None
    enc_out = torch.randn(B, 1 + K, d, device="cuda", generator=g)
    pos = torch.randint(0, rows, (B,), device="cuda", generator=g)
    neg = torch.randint(0, rows, (B, K), device="cuda", generator=g)
    scale = 1.0 / B
    rows_l, cp, cn, g_enc = ops.tem_loss(enc_out, table, pos, neg, bias=bias, pos_weight=pos_weight, grad_scale=scale)
    a = enc_out[:, 0].contiguous()
    b = enc_out[:, 1:].reshape(B * K, d).contiguous()
    loss, cp0, cn0, ga, gb = ops.ns_loss(a, table, pos.view(B, 1), neg.view(B, 1, K), anchor_b=b, bias=bias,
                                         pos_weight=pos_weight)
    assert torch.equal(rows_l, loss)                              # same arithmetic, only the addressing differs
    torch.testing.assert_close(cp, cp0.view(-1) * scale, rtol=1e-6, atol=0)
    torch.testing.assert_close(cn, cn0.view(B, K) * scale, rtol=1e-6, atol=0)
    torch.testing.assert_close(g_enc[:, 0], ga * scale, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(g_enc[:, 1:].reshape(B * K, d), gb * scale, rtol=1e-5, atol=1e-9)
    # loss combination + running sums in one launch
    il = torch.rand(B + 3, device="cuda", generator=g)
    acc_ps = torch.full((), 2.0, device="cuda")
    acc_il = torch.full((), 3.0, device="cuda")
    total = ops.tem_loss_finish(rows_l, il, acc_ps, acc_il)
    ps_ref, il_ref = rows_l.double().mean(), il.double().mean()
    assert abs(float(total) - float(ps_ref + il_ref)) <= 1e-6 * abs(float(ps_ref + il_ref))
    assert abs(float(acc_ps) - 2.0 - float(ps_ref)) <= 1e-5 * abs(float(ps_ref)) + 1e-6
    assert abs(float(acc_il) - 3.0 - float(il_ref)) <= 1e-5
    total2 = ops.tem_loss_finish(rows_l, il, acc_ps, acc_il)
    assert torch.equal(total, total2)                             # fixed summation order


@pytest.mark.parametrize("n,rows,d", [(900, 300, 128), (12000, 18001, 128), (40000, 700, 64), (300000, 2_000_001, 128)])
def test_sort_then_reduce_equals_the_one_call_path_bitwise(ops, n, rows, d):
    """All three sort paths (one-CTA counting sort, one-CTA radix, multi-CTA radix): the split calls leave exactly
    what psb_scatter_reduce_rows leaves -- unique rows, count, reduced rows, bias sums, dense scatter."""
    g = torch.Generator(device="cuda").manual_seed(n)
    pad = rows - 1
    idx_a = torch.randint(0, rows, (n,), device="cuda", generator=g)
    idx_b = torch.randint(0, min(rows, 97), (n // 3,), device="cuda", generator=g)       # a hot range: long segments
    src_a = torch.randn(n, d, device="cuda", generator=g)
    anchors = torch.randn(max(n // 9, 1), d, device="cuda", generator=g)
    sc_b = torch.randn(n // 3, device="cuda", generator=g)
    contribs = [ops.make_contrib(idx_a, src_a), ops.make_contrib(idx_b, anchors, src_div=3, scale=sc_b, to_bias=True)]
    uniq0, red0, redb0, nu0 = ops.scatter_reduce(contribs, rows, d, pad, want_rows=True, want_bias=True, device="cuda")
    n0 = int(nu0)
    n_total = n + n // 3
    ws = torch.empty(ops.scatter_workspace_bytes(n_total, rows), dtype=torch.uint8, device="cuda")
    uniq = torch.empty(n_total, dtype=torch.int32, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.scatter_sort([idx_a, idx_b], rows, pad, ws, uniq, nu)
    assert int(nu) == n0 and torch.equal(uniq[:n0], uniq0[:n0])
    red, redb = ops.scatter_reduce_sorted(contribs, rows, d, pad, ws, uniq, nu, want_rows=True, want_bias=True)
    assert torch.equal(red[:n0], red0[:n0]) and torch.equal(redb[:n0], redb0[:n0])
    # the dense scatter variant (one reduce per sort: the reduce consumes the workspace, so sort again)
    ops.scatter_sort([idx_a, idx_b], rows, pad, ws, uniq, nu.zero_())
    dense = torch.zeros(rows, d, device="cuda")
    dense_b = torch.zeros(rows, device="cuda")
    ops.scatter_reduce_sorted(contribs, rows, d, pad, ws, uniq, nu, dense_grad=dense, dense_bias_grad=dense_b)
    ref = torch.zeros(rows, d, device="cuda")
    ref[uniq0[:n0].long()] = red0[:n0]
    ref_b = torch.zeros(rows, device="cuda")
    ref_b[uniq0[:n0].long()] = redb0[:n0]
    assert torch.equal(dense, ref) and torch.equal(dense_b, ref_b)
    assert not bool(dense[pad].any())


def test_sqnorm_over_dense_tensors_and_row_lists(ops):
    from prodsearch_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    d, rows, cap, nu = 128, 4000, 700, 531
    dense = [torch.randn(5000, device="cuda", generator=g), torch.randn(33, 7, device="cuda", generator=g)]
    lst = torch.randperm(rows, device="cuda", generator=g)[:cap].int()
    compact = torch.randn(cap, d, device="cuda", generator=g)            # entries past nu must not count
    cbias = torch.randn(cap, device="cuda", generator=g)
    byrow = torch.randn(rows, d, device="cuda", generator=g)              # dense buffer: only the listed rows count
    nu_dev = torch.tensor([nu], dtype=torch.int32, device="cuda")
    arr = (_lib.AdamTensor * 2)(*[_lib.AdamTensor(None, t.data_ptr(), None, None, t.numel()) for t in dense])
    R = (_lib.AdamRows * 2)()
    R[0].rows, R[0].grad, R[0].n_rows, R[0].cap, R[0].d, R[0].table_rows = lst.data_ptr(), compact.data_ptr(), nu_dev.data_ptr(), cap, d, rows
    R[0].bias_grad = cbias.data_ptr()
    R[1].rows, R[1].grad, R[1].n_rows, R[1].cap, R[1].d, R[1].table_rows = lst.data_ptr(), byrow.data_ptr(), nu_dev.data_ptr(), cap, d, rows
    R[1].grad_by_row = 1
    wb = int(lib.psb_adam_sparse_workspace_bytes(arr, 2, R, 2))
    ws = torch.empty(wb, dtype=torch.uint8, device="cuda")
    out = torch.zeros(1, device="cuda")
    _lib.check(lib.psb_grad_sqnorm_sparse(arr, 2, R, 2, out.data_ptr(), ws.data_ptr(), wb, _lib.stream_ptr()), "sqnorm")
    ref = sum(float((t.double() ** 2).sum()) for t in dense)
    ref += float((compact[:nu].double() ** 2).sum()) + float((cbias[:nu].double() ** 2).sum())
    ref += float((byrow[lst[:nu].long()].double() ** 2).sum())
    assert abs(float(out) - ref) <= 1e-5 * ref
    out2 = torch.zeros(1, device="cuda")
    _lib.check(lib.psb_grad_sqnorm_sparse(arr, 2, R, 2, out2.data_ptr(), ws.data_ptr(), wb, _lib.stream_ptr()), "sqnorm")
    assert torch.equal(out, out2)


@pytest.mark.parametrize("world", [1, 3])
def test_lazy_peer_fetch_equals_catchup_then_fetch(world):
    """A reader that fetches resting rows of a row-sparse shard adds the catch-up series on the fly; the owner's copy
    is not touched.  Reference: the same rows after psb_adam_rows_catchup on every owner, then the plain fetch."""
    from prodsearch_b200 import _lib, peer
    lib = _lib.load()
    torch.manual_seed(5)
    rows, d, steps = 1201, 128, 300     # some rows rest longer than the series' 198 terms
    full = torch.randn(rows, d, device="cuda")
    groups = peer.PeerGroup.simulate(world, "cuda")
    tabs = [peer.PeerShardedTable(rows, d, g, pad_idx=rows - 1, full=full, sparse=True) for g in groups]
    lr, b1, b2, eps = 5e-3, 0.9, 0.999, 1e-9
    step_dev = torch.tensor([steps], dtype=torch.int64, device="cuda")
    tau = torch.arange(4096, dtype=torch.float64, device="cuda").clamp_(min=1)
    hist = torch.stack([lr / (1 - b1 ** tau), 1 / torch.sqrt(1 - b2 ** tau)], dim=1).float().reshape(-1).contiguous()
    g = torch.Generator(device="cuda").manual_seed(6)
    for t in tabs:                                   # rows that rested since different steps; some fresh, some current
        t.exp_avg.copy_(torch.randn(t.local_rows, d, device="cuda", generator=g) * 1e-2)
        t.exp_avg_sq.copy_(torch.rand(t.local_rows, d, device="cuda", generator=g) * 1e-4 + 1e-8)
        t.last_step.copy_(torch.randint(0, steps + 1, (t.local_rows,), device="cuda", generator=g).int())
    ids = torch.randint(0, rows, (700,), device="cuda", generator=g)
    before = [t.weight.detach().clone() for t in tabs]
    t0 = tabs[0]
    P = t0.shard.ptr_array()
    M, V, L = t0._m_buf.ptr_array(), t0._v_buf.ptr_array(), t0._last_buf.ptr_array()
    out = torch.empty(ids.numel(), d, device="cuda")
    _lib.check(lib.psb_peer_gather_rows_lazy(P, M, V, L, world, rows, d, ids.data_ptr(), ids.numel(), out.data_ptr(), None,
                                             -1, 0, None, lr, b1, b2, eps, 0, 4000.0, step_dev.data_ptr(), hist.data_ptr(),
                                             4096, _lib.stream_ptr()), "lazy gather")
    for t, w0 in zip(tabs, before):
        assert torch.equal(t.weight.detach(), w0)                  # nothing written back
    # reference: bring every owner's rows up to date in place, then fetch plainly
    for t in tabs:
        R = _lib.AdamRows()
        R.p, R.m, R.v, R.last_step = t.weight.data_ptr(), t.exp_avg.data_ptr(), t.exp_avg_sq.data_ptr(), t.last_step.data_ptr()
        R.d, R.table_rows = d, t.local_rows
        _lib.check(lib.psb_adam_rows_catchup(R, None, None, 0, -1, lr, b1, b2, eps, 0, 4000.0, step_dev.data_ptr(),
                                             hist.data_ptr(), 4096, _lib.stream_ptr()), "catchup")
        assert bool((t.last_step >= steps).all())
    ref = torch.empty_like(out)
    _lib.check(lib.psb_peer_gather_rows(P, world, rows, d, ids.data_ptr(), ids.numel(), ref.data_ptr(), None, -1, 0, None,
                                        _lib.stream_ptr()), "gather")
    moved = (ref - full[ids]).abs().max().item()
    assert moved > 1e-3                                            # the catch-up really moved rows
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
