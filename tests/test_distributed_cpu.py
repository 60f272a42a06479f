"""CPU, world_size 2 over gloo: the N>1 host logic (shard ranges, padded all-gather of results, CSR gather for
locate).  The per-rank engine is replaced by the CPU oracle here — the sharding code does not care who
answers the local slice; on the GPU box the same functions wrap the CUDA engine (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges():
    import __graft_entry__ as ge

    ge.load_package()
    from sdsl_lite_b200 import distributed as D

    for n in (0, 1, 7, 8, 100, 12345):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == D.shard_sizes(n, world)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import __graft_entry__ as ge
    import cases
    import texts

    ge.load_package()
    from sdsl_lite_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = ge.load_oracle().Oracle()
        nbits = 200001
        w = cases.random_words(nbits, 3)
        bv = orc.bv(w, nbits)
        idx = cases.rank_queries(nbits, 5, 10001)  # odd: uneven shards
        full = D.sharded_query(lambda i: bv.rank(i, 1), [idx])
        ok1 = bool((full == bv.rank(idx, 1)).all())
        local = D.sharded_query(lambda i: bv.rank(i, 1), [idx], gather=False)
        lo, hi = D.shard_range(len(idx), rank, world)
        ok2 = bool((local == bv.rank(idx[lo:hi], 1)).all())
        # two-column query (wt.rank) and the CSR gather for locate
        t = dict(texts.text_catalogue(zero_free=True, large=False))["dna"]
        wt = orc.wt_huff(t)
        rng = np.random.default_rng(1)
        qi, qc = texts.wt_queries(t, rng, 3001)
        ok3 = bool((D.sharded_query(lambda i, c: wt.rank(i, c), [qi, qc]) == wt.rank(qi, qc)).all())
        csa = orc.csa(t)
        pats = [t[s : s + 6] for s in rng.integers(0, len(t) - 6, 41)] + [b"", b"zzzz"]
        po = ge.load_oracle()
        flat, off = po.csr_patterns(pats)
        want = csa.locate(flat, off)
        got = D.sharded_locate(None, lambda f, o: csa.locate(f, o), flat, off)
        ok4 = bool((got[0] == want[0]).all() and (got[1] == want[1]).all())
        q.put((rank, ok1, ok2, ok3, ok4))
    finally:
        dist.destroy_process_group()


def test_sharded_queries_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert all(r[1:]), r
