"""CPU, world_size 2, gloo: the host-side logic of the row-sharded item table.

The exchange plan, the index / row / gradient all-to-alls and the sharded top-k merge are
backend-agnostic (prodsearch_b200/sharded.py); here the device kernels are replaced by plain torch
CPU stand-ins (test infrastructure only) and the results are compared with the unsharded answer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_gather(w, ids):
    return w[ids]


def _cpu_fold(w, ids, grads):
    if w.grad is None:
        w.grad = torch.zeros_like(w)
    w.grad.index_add_(0, ids, grads)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from prodsearch_b200 import sharded
        g = torch.Generator().manual_seed(0)                      # same data on every rank
        P, d, k = 203, 16, 7
        full = torch.randn(P + 1, d, generator=g)
        full[P] = 0
        table = sharded.ShardedTable.from_full(full, None, _cpu_gather, _cpu_fold, "cpu", pad_idx=P)
        assert table.weight.shape[0] == (P + 1 - rank + world - 1) // world
        # per-rank batch: different ids on each rank, duplicates, pad ids, ragged shapes
        gr = torch.Generator().manual_seed(100 + rank)
        tgt = torch.randint(0, P, (9,), generator=gr)
        neg = torch.randint(0, P, (9, 3), generator=gr)
        hist = torch.randint(0, P, (9, 5), generator=gr)
        hist[2, 3:] = P
        hist[4, :] = P
        neg[0, 0] = tgt[0]
        mini, (t2, n2, h2), pad = table.fetch([tgt, neg, hist])
        # bit-exact rows through the two all-to-alls and the remap
        assert torch.equal(mini[t2], full[tgt]) and torch.equal(mini[n2], full[neg]) and torch.equal(mini[h2], full[hist])
        assert pad >= 0 and torch.equal(mini[pad], full[P])
        assert mini.shape[0] == torch.unique(torch.cat([tgt, neg.reshape(-1), hist.reshape(-1)])).numel()
        # gradient push: every rank contributes a dense mini-gradient; owners fold them
        mg = torch.zeros_like(mini)
        up = torch.randn(9, d, generator=gr)
        mg.index_add_(0, t2, up)
        mg.index_add_(0, h2.reshape(-1), up.repeat_interleave(5, 0) * 0.5)
        mg[pad] = 0
        table.push_grads(mg.detach())
        # reference: the same contributions accumulated on the full table, all ranks summed
        ref = torch.zeros(P + 1, d)
        for r in range(world):
            g2 = torch.Generator().manual_seed(100 + r)
            tg = torch.randint(0, P, (9,), generator=g2)
            _ = torch.randint(0, P, (9, 3), generator=g2)
            hs = torch.randint(0, P, (9, 5), generator=g2)
            hs[2, 3:] = P
            hs[4, :] = P
            u2 = torch.randn(9, d, generator=g2)
            ref.index_add_(0, tg, u2)
            ref.index_add_(0, hs.reshape(-1), u2.repeat_interleave(5, 0) * 0.5)
        ref[P] = 0
        assert torch.allclose(table.weight.grad, ref[rank::world], atol=1e-6)
        # sharded catalog ranking == global ranking (lower id first), ragged query counts per rank
        q = torch.randint(-2, 3, (3 + rank, d), generator=gr).float()
        qtab = torch.randint(-2, 3, (P + 1, d), generator=g).float()      # integer data: exact ties
        qtab[10:20] = qtab[3]
        t2b = sharded.ShardedTable.from_full(qtab, None, _cpu_gather, _cpu_fold, "cpu", pad_idx=P)

        def topk(qa, w, kk, n_local, base, stride, bias):
            ids = base + stride * torch.arange(n_local)
            i, s_ = oracle.topk_lower_id_first((qa @ w[:n_local].t()).numpy(), kk, np.tile(ids.numpy(), (qa.shape[0], 1)))
            return torch.from_numpy(i), torch.from_numpy(s_)

        def merge(ids, sc):
            i, s_ = oracle.merge_shard_topk(list(ids.numpy()), list(sc.numpy()), k)
            return torch.from_numpy(i), torch.from_numpy(s_)
        ids, sc = sharded.sharded_rank_catalog(q, t2b, P, k, topk, merge)
        ref_i, ref_s = oracle.topk_lower_id_first((q @ qtab[:P].t()).numpy(), k)
        assert np.array_equal(ids.numpy(), ref_i) and np.array_equal(sc.numpy(), ref_s)
        # replicated-parameter gradient averaging
        lin = torch.nn.Linear(4, 3)
        with torch.no_grad():
            for p_ in lin.parameters():
                p_.fill_(1.0)
                p_.grad = torch.full_like(p_, float(rank + 1))
        sharded.DenseGradAllReduce(lin).reduce()
        assert all(torch.allclose(p_.grad, torch.full_like(p_, (1 + world) / 2)) for p_ in lin.parameters())
        out.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_table_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
