"""B=1 latency of one verification step with the hypothesis set sharded over the ranks (torchrun):
fused NVLink exchange (ahv_verify_sharded) against NCCL all-gather + merge.  AHV_N (50000), AHV_B (1)."""
import importlib, json, os, statistics, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
ahv = importlib.import_module("3dahv_b200")
N, B = int(os.environ.get("AHV_N", "50000")), int(os.environ.get("AHV_B", "1"))
W1, W2, b2, vs, vt, normals = bench.synthetic_inputs(torch, B, N)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R = ahv.ops.rotations_from_normals(normals.to(dev))
vs, vt = vs.to(dev), vt.to(dev)
peer = ahv.dist.PeerExchange(B, dev)
out = {}
ref = None
for name, sv in (("nccl_allgather_merge", ahv.dist.ShardedVerifier(v)), ("fused_peer_exchange", ahv.dist.ShardedVerifier(v, peer=peer))):
    for _ in range(10):
        val, idx, Rb = sv.score(vs, vt, R, k=1)
    dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    for a, b in ev:
        a.record(); val, idx, Rb = sv.score(vs, vt, R, k=1); b.record()
    torch.cuda.synchronize()
    t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"p50_us": float(t), "idx": idx[:, 0].tolist()[:4]}
    ref = idx if ref is None else ref
    assert torch.equal(ref, idx)
lo, hi = ahv.dist.shard_bounds(N, rank, world)
gv = ahv.GraphedVerifier(v, B, hi - lo, k=1, device=dev, peer=peer, idx_offset=lo)
gv(vs, vt, R[lo:hi])
for _ in range(10):
    gv()
dist.barrier(); torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
for a, b in ev:
    a.record(); o = gv(); b.record()
torch.cuda.synchronize()
t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
out["fused_peer_exchange_graph_replay"] = {"p50_us": float(t), "idx": o.topk_idx[:, 0].tolist()[:4]}
assert torch.equal(o.topk_idx, ref)
single = v.score(vs, vt, R, k=1, return_scores=False)
assert torch.equal(single.topk_idx, ref)
if rank == 0:
    print(json.dumps({"what": "sharded verification step latency (eager launches, max over ranks)", "pairs": B, "hyps": N, "gpus": world, **out}))
dist.barrier()
peer.close()
dist.destroy_process_group()
