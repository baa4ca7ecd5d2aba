"""Which property of a peer allocation makes P2P row loads slow?  (profiles/r02m: same kernel, same ids: 0.42 ms on one
4 GB allocation, 6.35 ms on another.)  torchrun, 2 ranks: a series of peer allocations of different sizes / orders,
each read with the same random global ids."""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from prodsearch_b200 import peer
    from prodsearch_b200._lib import check, load, stream_ptr
    pg = peer.PeerGroup(device="cuda")
    d, n = 128, 1_000_000
    out = torch.empty(n, d, device="cuda")
    res = []

    def probe(label, nbytes, rows_total, touch):
        buf = pg.alloc(nbytes)
        if touch == "normal":
            buf.local.view(torch.float32).normal_()
        elif touch == "param":
            w = buf.view(torch.float32, (nbytes // (d * 4), d))
            p = torch.nn.Parameter(w)
            with torch.no_grad():
                p.normal_()
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.Generator(device="cuda").manual_seed(5 + rank)
        ids = torch.randint(0, rows_total, (n,), device="cuda", generator=g)
        def run():
            check(load().psb_peer_gather_rows(buf.ptr_array(), world, rows_total, d, ids.data_ptr(), n, out.data_ptr(), None, -1, 0,
                                              None, stream_ptr()), "g")
        for _ in range(2):
            run()
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            run()
        e.record(); torch.cuda.synchronize()
        res.append({"label": label, "bytes": nbytes, "rows_total": rows_total, "ms": round(s.elapsed_time(e) / 5, 3),
                    "ptr_mod_2MB": buf.ptr % (2 << 20), "peer_ptr_mod_2MB": buf.ptrs[(rank + 1) % world] % (2 << 20)})
        return buf
    keep = []
    keep.append(probe("A 8M rows, first", 8_000_000 * 512, 16_000_000, "normal"))
    keep.append(probe("B 8M rows, second", 8_000_000 * 512, 16_000_000, "normal"))
    keep.append(probe("C 8M+1 rows", 8_000_001 * 512, 16_000_001, "normal"))
    keep.append(probe("D 8M+1 rows, rows_total even", 8_000_001 * 512, 16_000_000, "normal"))
    keep.append(probe("E 8M rows, rows_total odd", 8_000_000 * 512, 15_999_999, "normal"))
    keep.append(probe("F 8M+1 rows via Parameter", 8_000_001 * 512, 16_000_001, "param"))
    keep.append(probe("G 8M+4096 rows", 8_004_096 * 512, 16_008_192, "normal"))
    if rank == 0:
        for r in res:
            print(json.dumps(r), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
