"""Soak test of the NVLink exchange protocol (torchrun): thousands of consecutive sharded steps on one exchange
buffer - fused k = 1 exchange, top-k exchange kernel, B and k changing from step to step, eager and CUDA-graph
replays - every result compared with the unsharded answer computed beforehand.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/soak_sharded.py [steps]
"""
import importlib, json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
ahv = importlib.import_module("3dahv_b200")
W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, 8, 4000)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R = ahv.ops.rotations_from_normals(normals.to(dev))
vs, vt = vs.to(dev), vt.to(dev)
peer = ahv.dist.PeerExchange(8, dev, max_k=8)
sv = ahv.dist.ShardedVerifier(v, peer=peer)
cases = [(1, 4000, 1), (8, 1000, 1), (3, 2500, 8), (1, 64, 4), (5, 4000, 1), (2, 17, 8), (8, 4000, 2)]
want = []
for B, N, k in cases:          # the same on every rank (deterministic kernels)
    r = v.score(vs[:B], vt[:B], R[:N], k=k, return_scores=False)
    want.append((r.topk_val.clone(), r.topk_idx.clone(), r.R_best.clone()))
lo, hi = ahv.dist.shard_bounds(4000, rank, world)
gv = ahv.GraphedVerifier(v, 8, hi - lo, k=1, device=dev, peer=peer, idx_offset=lo)
gv(vs, vt, R[lo:hi])
g_want = v.score(vs, vt, R, k=1, return_scores=False)
bad = torch.zeros(1, device=dev, dtype=torch.int64)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    c = (i * 5 + i // 7) % len(cases)
    B, N, k = cases[c]
    if i % 3 == 2:                                    # a graph replay in between
        o = gv()
        ok = torch.equal(o.topk_idx, g_want.topk_idx) & torch.equal(o.topk_val, g_want.topk_val) & torch.equal(o.R_best, g_want.R_best)
        bad += 0 if ok else 1
        continue
    val, idx, Rb = sv.score(vs[:B], vt[:B], R[:N], k=k)
    kk = min(k, N)
    ok = torch.equal(idx[:, :kk], want[c][1]) and torch.equal(val[:, :kk], want[c][0]) and torch.equal(Rb[:, :kk], want[c][2])
    bad += 0 if ok else 1
torch.cuda.synchronize()
dt = time.perf_counter() - t0
dist.all_reduce(bad)
done = peer.check()
if rank == 0:
    print(json.dumps({"what": "soak of the NVLink exchange: alternating fused k=1 / top-k exchange / graph replays, changing B and k",
                      "gpus": world, "steps": steps, "exchanges_completed": done, "mismatching_steps_all_ranks": int(bad),
                      "seconds": dt, "steps_per_s": steps / dt}), flush=True)
dist.barrier()
peer.close()
dist.destroy_process_group()
