"""World-size-2 CPU test (gloo) of the multi-GPU sharding logic used by bench.py: terms are dealt round-robin
along the CKD-term axis, every rank keeps the global group numbering and forms AIK-weighted partial sums, one
reduce(sum) to rank 0 must reproduce the sequential SOS_AGGREGATE chain (SOS_AGGREGATE.F:397-413, 467-488).
The per-term results come from the oracle here (no GPU); on the GPU box the same layout is filled by libsosgpu."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as orc
    from util import oracle_term
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    syn = pkg.synth
    wl = syn.config_ckd_band(npoints=3, seed=7, nb_gauss=6, os_nb=12, max_terms=5)
    ngroup, o = len(wl.optics), wl.optics[0]
    W, rs = 2 * o.nbmu + 1, o.os_nb + 1
    part = np.zeros((ngroup, rs, 3, W))
    sc = np.zeros((ngroup, 5))                       # emoins, eplus, sum a*exp(-ttot_tronc / vrai / tauout)
    for i in range(rank, len(wl.terms), world):      # this rank's shard
        t = wl.terms[i]
        r = oracle_term(orc, wl.optics[t.optics], t)
        part[t.optics, :r.n_fourier] += t.aik * r.rec
        sc[t.optics] += t.aik * np.array([r.emoins, r.eplus, np.exp(-r.ttot_tronc), np.exp(-r.ttot_vrai), np.exp(-r.tauout)])
    tp, ts = torch.from_numpy(part), torch.from_numpy(sc)
    dist.reduce(tp, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(ts, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        ok = True
        for g in range(ngroup):                      # sequential reference chain
            agg = orc.Aggregate(o.nbmu, rs)
            for t in wl.terms:
                if t.optics == g:
                    agg.add(t.aik, oracle_term(orc, wl.optics[g], t))
            n = min(agg.nres, rs)                    # records beyond the longest series are the reference's zero padding
            ok &= np.allclose(tp[g, :n].numpy(), agg.res[:n], rtol=1e-12, atol=1e-300)
            ok &= bool(np.all(agg.res[n:agg.nres] == 0.0))
            ok &= np.isclose(float(ts[g, 0]), agg.sc["emoins"], rtol=1e-12) and np.isclose(float(ts[g, 1]), agg.sc["eplus"], rtol=1e-12)
            ok &= np.isclose(-np.log(float(ts[g, 2])), agg.sc["ttot_tronc"], rtol=1e-12)
            ok &= np.isclose(-np.log(float(ts[g, 4])), agg.sc["tauout"], rtol=1e-12, atol=1e-15)
        q.put(bool(ok))
    dist.destroy_process_group()


def test_round_robin_shards_reduce_to_sequential_aggregate():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _worker_wavelengths(rank, world, port, q):
    """Wavelength-sharded layout of bench.py (default): whole spectral points per rank by LPT, local group numbering,
    results gathered on rank 0 in rank order; the gathered list must be a permutation of the band that bench.py's
    bookkeeping (groups_of_rank, rank-order concatenation) maps back to the right spectral points."""
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from oracle import oracle as orc
    from util import oracle_term
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("radiativetransfer-sos_b200")
    wl = pkg.synth.config_ckd_band(npoints=5, seed=11, nb_gauss=6, os_nb=12, max_terms=5)
    part = bench.lpt_partition(bench.point_costs(wl), world)
    my_points = [p for p in range(len(wl.optics)) if part[p] == rank]
    o = wl.optics[0]
    W, rs = 2 * o.nbmu + 1, o.os_nb + 1
    mine = np.zeros((len(my_points), rs, 3, W))
    for g, p in enumerate(my_points):                # whole wavelengths: the local chain IS the reference chain
        agg = orc.Aggregate(o.nbmu, rs)
        for t in wl.terms:
            if t.optics == p:
                agg.add(t.aik, oracle_term(orc, wl.optics[p], t))
        mine[g] = agg.res[:rs]
    groups_of_rank = [int(np.sum(part == r)) for r in range(world)]
    # ragged gather = send/recv to the root, as sosgpu_batch_gather_tables does with ncclSend/ncclRecv
    if rank == 0:
        gathered = [torch.from_numpy(mine)] + [torch.zeros((n, rs, 3, W), dtype=torch.float64) for n in groups_of_rank[1:]]
        for r in range(1, world):
            if groups_of_rank[r]:
                dist.recv(gathered[r], src=r)
    elif len(my_points):
        dist.send(torch.from_numpy(mine), dst=0)
    if rank == 0:
        allg = torch.cat(gathered).numpy()
        order = [p for r in range(world) for p in range(len(wl.optics)) if part[p] == r]   # rank-order concatenation
        ok = sorted(order) == list(range(len(wl.optics))) and sum(groups_of_rank) == len(wl.optics)
        for k, p in enumerate(order):
            agg = orc.Aggregate(o.nbmu, rs)
            for t in wl.terms:
                if t.optics == p:
                    agg.add(t.aik, oracle_term(orc, wl.optics[p], t))
            ok &= bool(np.array_equal(allg[k], agg.res[:rs]))
        # the partition is balanced: no rank carries more than the ideal share plus one point's cost
        c = bench.point_costs(wl)
        loads = [c[part == r].sum() for r in range(world)]
        ok &= max(loads) <= c.sum() / world + c.max() + 1e-9
        q.put(bool(ok))
    dist.destroy_process_group()


def test_wavelength_sharding_lpt_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_wavelengths, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_lpt_partition_balances_the_bench_band():
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.make_workload(96 * 2)
    c = bench.point_costs(wl)
    for world in (2, 4, 8):
        part = bench.lpt_partition(c, world)
        loads = np.array([c[part == r].sum() for r in range(world)])
        assert loads.max() / loads.mean() < 1.02, (world, loads)
