"""CPU, world_size 2 over gloo: the host logic of hypothesis sharding (bounds,
padding, all-gather of per-rank top-k, deterministic merge) with the oracle as
the per-rank scorer and a torch merge standing in for the CUDA merge kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _merge_torch(vals, idxs):
    """Reference merge: score desc, ties -> lowest global index, -1 ignored."""
    P, B, k = vals.shape
    v = vals.permute(1, 0, 2).reshape(B, P * k).double()
    i = idxs.permute(1, 0, 2).reshape(B, P * k)
    v = torch.where(i < 0, torch.full_like(v, -float("inf")), v)
    order = np.lexsort((i.numpy(), -v.numpy()), axis=-1)[:, :k]
    order = torch.from_numpy(order)
    return torch.gather(v, 1, order).float(), torch.gather(i, 1, order)


def _worker(rank, world, port, N, k, out_q):
    sys.path.insert(0, ROOT)
    import importlib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ahv = importlib.import_module("3dahv_b200")
    from oracle import ahv_oracle as orc

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "shared_n3000_b3.npz")))
    w = dict(np.load(os.path.join(ROOT, "tests", "golden", "weights.npz")))
    vs, vt, R = torch.from_numpy(g["vol_src"]), torch.from_numpy(g["vol_tgt"]), torch.from_numpy(g["R"][:N])

    def score_fn(vol_src, vol_tgt, R_slice, kk, idx_offset):
        s = orc.score_c(vol_src.numpy(), vol_tgt.numpy(), R_slice.numpy(), w["W1"], w["W2"], w["b2"], nthreads=2)
        val, idx = orc.select_np(s, kk)
        return torch.from_numpy(val), torch.from_numpy(idx + idx_offset)

    sv = ahv.dist.ShardedVerifier(None, merge=_merge_torch, score_fn=score_fn)
    val, idx, Rb = sv.score(vs, vt, R, k=k)
    out_q.put((rank, val.numpy(), idx.numpy(), Rb.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,k", [(301, 4), (3, 4)])
def test_sharded_topk_equals_single_process(golden, oracle, N, k):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + N
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = golden["shared_n3000_b3"]
    ref_val, ref_idx = oracle.select_np(g["scores"][:, :N], k)
    kk = min(k, N)
    for rank, val, idx, Rb in res:
        assert np.array_equal(idx, ref_idx), (rank, idx, ref_idx)
        np.testing.assert_allclose(val, ref_val, rtol=3e-6)
        assert np.array_equal(Rb, g["R"][:N][ref_idx])
        assert val.shape == (3, kk)
    # every rank holds bit-identical results
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


def test_shard_bounds(ahv):
    sb = ahv.dist.shard_bounds
    for N in (0, 1, 7, 8, 50000, 50001):
        for P in (1, 2, 4, 8):
            spans = [sb(N, r, P) for r in range(P)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi >= lo for lo, hi in spans)
    assert sb(50000, 7, 8) == (43750, 50000)
