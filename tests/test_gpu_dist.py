"""GPU, >=2 devices: hypothesis sharding over NCCL gives the single-GPU result on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, k, out_q):
    sys.path.insert(0, ROOT)
    import importlib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ahv = importlib.import_module("3dahv_b200")
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "shared_n3000_b3.npz")))
    w = dict(np.load(os.path.join(ROOT, "tests", "golden", "weights.npz")))
    T = lambda a: torch.from_numpy(a).to(dev)
    v = ahv.HypothesisVerifier(T(w["W1"]), T(w["W2"]), T(w["b2"]))
    sv = ahv.dist.ShardedVerifier(v)
    val, idx, Rb = sv.score(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"]), k=k)
    single = v.score(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"]), k=k, return_scores=False)
    # fused exchange: the scoring kernels trade their winners through NVLink peer memory (no NCCL in the step)
    peer = ahv.dist.PeerExchange(8, dev)
    fv = ahv.dist.ShardedVerifier(v, peer=peer)
    fused = []
    for rep in range(3):                                  # repeated exchanges exercise both parities
        fval, fidx, fRb = fv.score(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"]), k=1)
        fused.append((fval.cpu().numpy(), fidx.cpu().numpy(), fRb.cpu().numpy()))
    # the sharded step is NCCL-free, hence CUDA-graph capturable: replay it and compare again
    lo, hi = ahv.dist.shard_bounds(3000, rank, world)
    gv = ahv.GraphedVerifier(v, 3, hi - lo, k=1, device=dev, peer=peer, idx_offset=lo)
    for rep in range(3):
        out = gv(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"][lo:hi]))
        fused.append((out.topk_val.cpu().numpy(), out.topk_idx.cpu().numpy(), out.R_best.cpu().numpy()))
    Rp = torch.stack([T(g["R"][b * 900:(b + 1) * 900]) for b in range(3)]).contiguous()   # per-pair sets [3,900,3,3]
    pval, pidx, pRb = fv.score(T(g["vol_src"]), T(g["vol_tgt"]), Rp, k=1)
    psingle = v.score(T(g["vol_src"]), T(g["vol_tgt"]), Rp, k=1, return_scores=False)
    torch.cuda.synchronize()
    out_q.put((rank, val.cpu().numpy(), idx.cpu().numpy(), Rb.cpu().numpy(), single.topk_val.cpu().numpy(),
               single.topk_idx.cpu().numpy(), fused,
               (pval.cpu().numpy(), pidx.cpu().numpy(), pRb.cpu().numpy(), psingle.topk_val.cpu().numpy(),
                psingle.topk_idx.cpu().numpy(), psingle.R_best.cpu().numpy())))
    assert peer.check() == 3 + (2 + 3) + 1               # eager x3, graphed: 2 warm-ups + 3 replays, per-pair x1; none timed out
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_single_gpu_nccl(golden):
    world, k = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, val, idx, Rb, sval, sidx, fused, pp in res:
        assert np.array_equal(idx, sidx) and np.array_equal(val, sval)
        assert np.array_equal(Rb, golden["shared_n3000_b3"]["R"][idx])
        for fval, fidx, fRb in fused:                     # fused peer exchange == unsharded top-1, on every rank
            assert np.array_equal(fidx[:, 0], sidx[:, 0]) and np.array_equal(fval[:, 0], sval[:, 0])
            assert np.array_equal(fRb[:, 0], golden["shared_n3000_b3"]["R"][sidx[:, 0]])
        pval, pidx, pRb, sv, si, sR = pp
        assert np.array_equal(pidx, si) and np.array_equal(pval, sv) and np.array_equal(pRb, sR)
    assert np.array_equal(res[0][2], res[1][2])
