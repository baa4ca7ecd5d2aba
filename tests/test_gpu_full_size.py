"""GPU: the hot-path kernels at BASELINE.json's FULL size -- a 16M x 128 fp32 item table (8.2 GB, configs[4]) --
checked through size-independent properties (the CPU oracle cannot finish at this size): bit-exact gathers against
plain torch indexing, gather -> scatter round trips, exact small-integer multiplicities, and the full-catalog
top-100 of every mode against a chunked plain-PyTorch fp32 scoring of the whole table."""
import pytest
import torch

pytestmark = pytest.mark.gpu

ROWS, D = 16_000_000, 128


@pytest.fixture(scope="module")
def table():
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~40 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(666)
    t = torch.empty(ROWS + 1, D, device="cuda")
    t.normal_(generator=g)
    t[ROWS] = 0                                           # pad row (item_transformer.py:46-49)
    yield t
    del t
    torch.cuda.empty_cache()


@pytest.fixture(scope="module")
def ops():
    from prodsearch_b200 import ops as _ops
    return _ops


def test_gather_16m_bit_exact(ops, table):
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, ROWS + 1, (1_000_000,), generator=g, device="cuda")
    out = ops.gather_rows(table, idx)
    assert torch.equal(out, table[idx])
    pools = idx[:400_000].view(40_000, 10).contiguous()
    mean, _, _ = ops.gather_meanpool(table, pools, pad_idx=ROWS)
    # masked mean in token order, (x * m) rounded before the add as the reference does (text_encoder.py:8)
    m = pools.ne(ROWS)
    acc = torch.zeros(40_000, D, device="cuda")
    for j in range(10):
        acc = acc + table[pools[:, j]] * m[:, j:j + 1].float()
    ref = acc / m.sum(1, keepdim=True).clamp(min=1).float()
    assert torch.equal(mean, ref)


def test_gather_scatter_round_trip_16m(ops, table):
    """Every row id once: the reduced rows ARE the gathered rows, bit for bit, and come back sorted; three copies
    of every id with scales (1, 1, 1): (x + x) + x exactly."""
    g = torch.Generator(device="cuda").manual_seed(2)
    idx = torch.randperm(ROWS, generator=g, device="cuda")[:2_000_000].contiguous()
    src = ops.gather_rows(table, idx)
    uniq, red, _, nu = ops.scatter_reduce([ops.make_contrib(idx, src)], ROWS + 1, D, drop_idx=ROWS)
    n = int(nu.item())
    assert n == idx.numel()
    assert torch.equal(uniq[:n].long(), idx.sort().values)
    assert torch.equal(red[:n], table[uniq[:n].long()])
    idx3 = torch.cat([idx[:300_000]] * 3)
    src3 = torch.cat([src[:300_000]] * 3)
    uniq3, red3, _, nu3 = ops.scatter_reduce([ops.make_contrib(idx3, src3)], ROWS + 1, D, drop_idx=ROWS)
    n3 = int(nu3.item())
    assert n3 == 300_000
    rows = table[uniq3[:n3].long()]
    assert torch.equal(red3[:n3], (rows + rows) + rows)
    # idempotence: the same call returns the same bits
    uniq4, red4, _, _ = ops.scatter_reduce([ops.make_contrib(idx3, src3)], ROWS + 1, D, drop_idx=ROWS)
    assert torch.equal(red3[:n3], red4[:n3]) and torch.equal(uniq3[:n3], uniq4[:n3])


def test_ns_loss_16m_matches_torch(ops, table):
    g = torch.Generator(device="cuda").manual_seed(3)
    n, k = 100_000, 5
    anchor = torch.randn(n, D, generator=g, device="cuda") * 0.1
    pos = torch.randint(0, ROWS, (n, 1), generator=g, device="cuda")
    neg = torch.randint(0, ROWS, (n, 1, k), generator=g, device="cuda")
    loss, cp, cn, ga, _ = ops.ns_loss(anchor, table, pos, neg)
    rows = torch.cat([table[pos.view(n, 1)], table[neg.view(n, k)]], 1)              # [n, 1+k, d]
    x = torch.einsum("nd,nkd->nk", anchor.double(), rows.double())
    t = torch.zeros_like(x)
    t[:, 0] = 1
    ref = (x.clamp(min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).sum(1)
    assert torch.allclose(loss.double(), ref, rtol=1e-5, atol=1e-6)
    coef = torch.sigmoid(x) - t
    assert torch.allclose(torch.cat([cp.view(n, 1), cn.view(n, k)], 1).double(), coef, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ga.double(), torch.einsum("nk,nkd->nd", coef, rows.double()), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("m", [24, 384])
def test_catalog_top100_16m_all_modes(ops, table, m):
    from prodsearch_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(10 + m)
    q = torch.randn(m, D, generator=g, device="cuda")
    k = 100
    # plain-PyTorch fp32 reference: chunked scoring of the whole table, running top-k (ties -> lower id)
    best_s = torch.full((m, k), -float("inf"), device="cuda")
    best_i = torch.zeros((m, k), dtype=torch.int64, device="cuda")
    for c0 in range(0, ROWS, 1_000_000):
        s = q @ table[c0:c0 + 1_000_000].t()
        cs, ci = s.topk(k, dim=1)
        alls = torch.cat([best_s, cs], 1)
        alli = torch.cat([best_i, ci + c0], 1)
        order = torch.sort(alli, dim=1, stable=True).indices              # ids ascending, then stable by score
        alls, alli = alls.gather(1, order), alli.gather(1, order)
        order = torch.sort(alls, dim=1, descending=True, stable=True).indices
        best_s, best_i = alls.gather(1, order)[:, :k], alli.gather(1, order)[:, :k]
    prep = ops.catalog_prepare_f16(table, ROWS)
    assert prep.fits
    norm = ops.table_max_row_sqnorm(table, ROWS)
    out = {}
    for name, mode in (("exact", _lib.TOPK_EXACT), ("tf32", _lib.TOPK_TC), ("f16", _lib.TOPK_TC16)):
        ids, sc = ops.catalog_topk(q, table, k, n_items=ROWS, mode=mode, max_row_sqnorm=norm,
                                   prepared=prep if mode == _lib.TOPK_TC16 else None)
        out[name] = (ids, sc)
        assert bool((sc[:, :-1] >= sc[:, 1:]).all())                       # sorted
        assert int(ids.min()) >= 0 and int(ids.max()) < ROWS
        # the scores are the fp32 dot products of the returned ids (fp64 recomputation)
        ref_sc = torch.einsum("md,mkd->mk", q.double(), table[ids].double())
        assert torch.allclose(sc.double(), ref_sc, rtol=1e-5, atol=1e-4)
    # every tensor-core mode returns exactly the exact mode's lists
    for name in ("tf32", "f16"):
        assert torch.equal(out[name][0], out["exact"][0]) and torch.equal(out[name][1], out["exact"][1]), name
    # against the torch reference: scores within tolerance everywhere, ids equal wherever the neighbouring gaps
    # of the reference list exceed the summation-order noise
    ids, sc = out["exact"]
    assert torch.allclose(sc, best_s, rtol=1e-5, atol=2e-4)
    gap = (best_s[:, :-1] - best_s[:, 1:]).abs()
    safe = torch.ones_like(best_s, dtype=torch.bool)
    safe[:, :-1] &= gap > 1e-3
    safe[:, 1:] &= gap > 1e-3
    safe[:, -1] = False                                                    # the k-th entry competes with rank k+1
    assert float(safe.float().mean()) > 0.8
    assert torch.equal(ids[safe], best_i[safe])
