"""One rank of tests/test_group_gpu.py::test_group_one_process_per_gpu_torchrun: a group built with
sdslgpu_group_create_rank over an id broadcast by torch.distributed; NCCL and fused gathers, replicate, and
distributed.sharded_query under the NCCL backend — all against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402
import texts  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    orc = ge.load_oracle().Oracle()
    from sdsl_lite_b200 import distributed as D

    ids = [pkg.group_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    g = pkg.Group.create_rank(ids[0], world, rank, local)
    assert g.nranks == world and g.nlocal == 1 and g.first_rank == rank
    dev = torch.device("cuda", local)

    def to_dev(a):
        return torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)

    def host(t):
        a = t.cpu().numpy()
        return a.view(np.uint64) if a.dtype == np.int64 else a

    nbits, nq = 5_000_011, 200_003
    w = cases.random_words(nbits, 11)
    ob = orc.bv(w, nbits)
    rng = np.random.default_rng(5)  # same seed on every rank: the batch is identical everywhere
    idx = rng.integers(0, nbits + 1, nq, dtype=np.uint64)
    want = ob.rank(idx, 1)
    # rank 0 builds, everybody receives a replica
    src = pkg.BitVector(w, nbits, device=local) if rank == 0 else None
    bv = g.replicate(src, root=0)[0]
    m1 = bv.arg_count(1)
    sel = rng.integers(1, m1 + 1, nq, dtype=np.uint64)
    want_s = ob.select(sel, 1)
    d_idx, d_sel = to_dev(idx), to_dev(sel)
    sym = g.alloc(nq * 8)
    out = sym.tensor(0)
    gathers = [pkg.GATHER_NCCL] + ([pkg.GATHER_FUSED, pkg.GATHER_PACKED, pkg.GATHER_AUTO] if g.fused_possible else [])
    for order in (pkg.ORDER_BINNED, pkg.ORDER_DIRECT):
        bv.set_batch_order(order)
        for gather in gathers:
            out.fill_(-7)
            g.rank([bv], 1, [d_idx], [out], gather=gather)
            assert (host(out) == want).all(), ("rank", order, gather)
            out.fill_(-7)
            g.select([bv], 1, [d_sel], [out], gather=gather)
            assert (host(out) == want_s).all(), ("select", order, gather)
            # asynchronous form on torch's current stream
            out.fill_(-7)
            g.rank([bv], 1, [d_idx], [out], gather=gather, streams=[torch.cuda.current_stream()])
            torch.cuda.synchronize()
            assert (host(out) == want).all(), ("rank async", order, gather)
    # the torch.distributed form of the same thing (NCCL all_gather of CUDA tensors, and of host results)
    full = D.sharded_query(lambda q: bv.rank(q, 1), [d_idx])
    assert full.is_cuda and (host(full) == want).all()
    full = D.sharded_query(lambda q: bv.rank(q, 1), [idx])
    assert isinstance(full, np.ndarray) and (full == want).all()
    t = dict(texts.text_catalogue(zero_free=True, large=False))["dna"]
    csa = pkg.CsaWt(t, device=local)
    pats = [t[s : s + 6] for s in rng.integers(0, len(t) - 6, 5001)] + [b"", b"zzzz"]
    flat, off = pkg.csr_patterns(pats)
    oc = orc.csa(t)
    cnt_out = sym.tensor(0)[: len(pats)]
    for gather in gathers:
        cnt_out.fill_(-7)
        g.fm_count([csa], [to_dev(flat)], [to_dev(off)], [cnt_out], gather=gather)
        assert (host(cnt_out) == oc.count(flat, off)).all(), ("fm_count", gather)
    got = D.sharded_locate(None, lambda f, o: csa.locate(f, o), flat, off)
    wl = oc.locate(flat, off)
    assert (got[0] == wl[0]).all() and (got[1] == wl[1]).all()
    sym.release()
    csa.close()
    bv.close()
    if src is not None:
        src.close()
    g.close()
    dist.barrier()
    dist.destroy_process_group()
    print("group worker ok", rank, flush=True)


if __name__ == "__main__":
    main()
