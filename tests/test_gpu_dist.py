"""GPU, >=2 devices: a hypothesis set sharded over ranks gives the single-GPU result on every rank - through NCCL
all-gather + merge, through the NVLink peer exchange fused into the scoring kernel (k = 1), through the
top-k exchange kernel (k > 1), with B and k changing from step to step on one exchange buffer, with empty
shards, under CUDA-graph replay and through the host-buffer entry."""
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
    cpu = lambda *ts: tuple(t.cpu().numpy() for t in ts)
    v = ahv.HypothesisVerifier(T(w["W1"]), T(w["W2"]), T(w["b2"]))
    vs, vt, R = T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"])
    out = {}
    # (1) NCCL path: all-gather of the per-rank top-k + merge kernel
    sv = ahv.dist.ShardedVerifier(v)
    out["nccl"] = cpu(*sv.score(vs, vt, R, k=k))
    single = v.score(vs, vt, R, k=k, return_scores=False)
    out["single"] = cpu(single.topk_val, single.topk_idx, single.R_best)
    # (2) peer exchange: k = 1 fused into the scoring kernel, k > 1 through the exchange kernel, one buffer
    peer = ahv.dist.PeerExchange(8, dev, max_k=k)
    fv = ahv.dist.ShardedVerifier(v, peer=peer, check_every=2)
    steps = 0
    out["fused_k1"] = []
    for rep in range(3):                                  # repeated exchanges exercise both parities
        out["fused_k1"].append(cpu(*fv.score(vs, vt, R, k=1)))
        steps += 1
    out["fused_topk"] = cpu(*fv.score(vs, vt, R, k=k))
    steps += 1
    # (3) B and k alternate between consecutive exchanges on the same buffer (fixed-capacity entry stride)
    out["alternate"] = []
    for rep in range(6):
        Bv, kv = (1, 1) if rep % 2 == 0 else (3, min(k, 5))
        r = fv.score(vs[:Bv], vt[:Bv], R[: 1000 + 100 * rep], k=kv)
        s1 = v.score(vs[:Bv], vt[:Bv], R[: 1000 + 100 * rep], k=kv, return_scores=False)
        out["alternate"].append((cpu(*r), cpu(s1.topk_val, s1.topk_idx, s1.R_best)))
        steps += 1
    # (4) fewer hypotheses than ranks*k, and an empty shard on the last rank
    out["tiny"] = []
    for n_tiny, kv in ((1, 1), (1, 4), (3, 4), (5, 2)):
        r = fv.score(vs, vt, R[:n_tiny], k=kv)
        s1 = v.score(vs, vt, R[:n_tiny], k=kv, return_scores=False)
        out["tiny"].append((n_tiny, kv, cpu(*r), cpu(s1.topk_val, s1.topk_idx, s1.R_best)))
        steps += 1
    # (5) the sharded step is NCCL-free, hence CUDA-graph capturable: replay it and compare again (k = 1 and k > 1)
    lo, hi = ahv.dist.shard_bounds(3000, rank, world)
    out["graph"] = []
    for kg in (1, k):
        gv = ahv.GraphedVerifier(v, 3, hi - lo, k=kg, device=dev, peer=peer, idx_offset=lo)
        steps += 2                                        # warm-ups inside the constructor
        for rep in range(3):
            o = gv(vs, vt, R[lo:hi])
            out["graph"].append((kg, cpu(o.topk_val, o.topk_idx, o.R_best)))
            steps += 1
    # (6) per-pair rotation sets, bf16 volumes
    Rp = torch.stack([R[b * 900:(b + 1) * 900] for b in range(3)]).contiguous()   # [3,900,3,3]
    ps = v.score(vs, vt, Rp, k=1, return_scores=False)
    out["per_pair"] = (cpu(*fv.score(vs, vt, Rp, k=1)), cpu(ps.topk_val, ps.topk_idx, ps.R_best))
    bs = v.score(vs.bfloat16(), vt, R, k=k, return_scores=False)
    out["bf16"] = (cpu(*fv.score(vs.bfloat16(), vt, R, k=k)), cpu(bs.topk_val, bs.topk_idx, bs.R_best))
    steps += 2
    # (6b) two-pass selection (config 4) with both passes sharded: top-k exchange, then fused exchange
    rR, rv, (fval_, fi, fR_), cand = fv.refine(vs, vt, R, k=k, m=16, max_angle_deg=4.0, seed=3)
    sR, sv_, sfirst, scand = v.refine(vs, vt, R, k=k, m=16, max_angle_deg=4.0, seed=3)
    out["refine"] = (cpu(rR, rv, fi, cand), cpu(sR, sv_, sfirst.topk_idx, scand))
    steps += 2
    # (6c) the full score matrix of the sharded set (NCCL all-gather of the per-rank pieces)
    out["scores"] = (cpu(sv.scores(vs, vt, R[:2999])), cpu(v.score(vs, vt, R[:2999], k=1, return_scores=True).scores))
    # (7) host-buffer entry, sharded: host slices in, whole-set selection out
    sess = ahv.ops.HostSession(dev)
    Rh = g["R"][lo:hi]
    for kh in (1, k):
        _, hv, hi_, hR = ahv.ops.predict_host(torch.from_numpy(g["vol_src"]), torch.from_numpy(g["vol_tgt"]),
                                              torch.from_numpy(Rh), T(w["W1"]), T(w["W2"]), T(w["b2"]), k=kh, device=dev,
                                              session=sess, idx_offset=lo, peer=peer)
        out.setdefault("host", []).append((kh, (hv.numpy(), hi_.numpy(), hR.numpy())))
        steps += 1
    sess.close()
    torch.cuda.synchronize()
    assert peer.check() == steps                          # every exchange completed, none timed out
    out_q.put((rank, out))
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


def _same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_single_gpu(golden):
    world, k = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    Rg = golden["shared_n3000_b3"]["R"]
    for rank in range(world):
        o = res[rank]
        sval, sidx, sR = o["single"]
        assert np.array_equal(sR, Rg[sidx])
        assert _same(o["nccl"], o["single"])
        for f in o["fused_k1"]:                           # fused peer exchange == unsharded top-1, on every rank
            assert _same(f, (sval[:, :1], sidx[:, :1], sR[:, :1]))
        assert _same(o["fused_topk"], o["single"])        # top-k exchange kernel == unsharded top-k
        for got, want in o["alternate"]:
            assert _same(got, want)
        for n_tiny, kv, got, want in o["tiny"]:
            kk = min(kv, n_tiny)                          # the single-GPU call clamps k to N; the exchange pads
            assert np.array_equal(got[0][:, :kk], want[0]) and np.array_equal(got[1][:, :kk], want[1])
            assert np.array_equal(got[2][:, :kk], want[2])
            assert np.all(got[1][:, kk:] == -1) and np.all(np.isneginf(got[0][:, kk:]))
        for kg, got in o["graph"]:
            assert _same(got, (sval[:, :kg], sidx[:, :kg], sR[:, :kg]))
        assert _same(*o["per_pair"]) and _same(*o["bf16"]) and _same(*o["refine"]) and _same(*o["scores"])
        for kh, got in o["host"]:
            assert _same(got, (sval[:, :kh], sidx[:, :kh], sR[:, :kh]))
    assert _same(res[0]["fused_topk"], res[1]["fused_topk"])
