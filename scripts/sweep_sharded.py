"""BASELINE config 5 across GPUs (torchrun): hypotheses x pairs sweep with the hypothesis set sharded over the ranks,
winners exchanged by the kernels over NVLink (k = 1 fused, k = 32 through the exchange kernel).  One JSON line per
point on rank 0: whole-set hyp*pairs/s (max over ranks), and whether the sharded result equals the unsharded one
(checked on rank 0 for every point whose unsharded run fits in a few hundred ms).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep_sharded.py [--quick]
"""
import argparse, importlib, json, os, statistics, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
ahv = importlib.import_module("3dahv_b200")
points = [(1, 10_000), (1, 100_000), (1, 1_000_000), (16, 100_000), (256, 10_000), (256, 100_000)]
if args.quick:
    points = [(1, 100_000), (16, 100_000), (256, 10_000)]
maxB, maxN = max(b for b, _ in points), max(n for _, n in points)
W1, W2, b2, vs_all, vt_all, normals = bench.synthetic_inputs(torch, maxB, maxN)
v = ahv.HypothesisVerifier(W1.to(dev), W2.to(dev), b2.to(dev))
R_all = ahv.ops.rotations_from_normals(normals.to(dev))
peer = ahv.dist.PeerExchange(maxB, dev, max_k=32)
sv = ahv.dist.ShardedVerifier(v, peer=peer)
for B, N in points:
    vs, vt, R = vs_all[:B].to(dev), vt_all[:B].to(dev), R_all[:N].contiguous()
    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        src = vs.to(dt)
        for k in (1, 32):
            if k == 32 and (B, N) not in ((1, 100_000), (256, 10_000)):
                continue
            for _ in range(3):
                val, idx, Rb = sv.score(src, vt, R, k=k)
            dist.barrier(); torch.cuda.synchronize()
            reps = max(3, min(30, int(2e8 * world // (B * N))))
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for a, b in ev:
                a.record(); val, idx, Rb = sv.score(src, vt, R, k=k); b.record()
            torch.cuda.synchronize()
            t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ok = None
            if rank == 0 and B * N <= 30_000_000:
                one = v.score(src, vt, R, k=k, return_scores=False)
                ok = bool(torch.equal(one.topk_idx, idx) and torch.equal(one.topk_val, val) and torch.equal(one.R_best, Rb))
            dist.barrier()
            if rank == 0:
                ms = float(t)
                print(json.dumps({"gpus": world, "pairs": B, "hyps": N, "vol": name, "k": k, "ms_p50": ms,
                                  "hyp_pairs_per_s": B * N / (ms * 1e-3), "sharded_equals_unsharded": ok}), flush=True)
# BASELINE config 4 sharded: 50 000-rotation grid + top-32 x 64 refinement, both passes sharded, no NCCL call
B4, N4, k4, m4 = 8, 50_000, 32, 64
vs4, vt4 = vs_all[:B4].to(dev), vt_all[:B4].to(dev)
grid4 = ahv.so3.grid_rotations(N4, device=dev)
for _ in range(3):
    r4 = sv.refine(vs4, vt4, grid4, k=k4, m=m4, max_angle_deg=5.0, seed=1)
dist.barrier(); torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    a.record(); r4 = sv.refine(vs4, vt4, grid4, k=k4, m=m4, max_angle_deg=5.0, seed=1); b.record()
torch.cuda.synchronize()
t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    one = v.refine(vs4, vt4, grid4, k=k4, m=m4, max_angle_deg=5.0, seed=1)
    ok = bool(torch.equal(one[0], r4[0]) and torch.equal(one[1], r4[1]) and torch.equal(one[2].topk_idx, r4[2][1]))
    print(json.dumps({"gpus": world, "config4_sharded_refine": True, "pairs": B4, "grid": N4, "k": k4, "m": m4, "ms_p50": float(t),
                      "hyp_pairs_per_s": B4 * (N4 + k4 * m4) / (float(t) * 1e-3), "sharded_equals_unsharded": ok}), flush=True)
dist.barrier()
peer.check()
dist.barrier()
peer.close()
dist.destroy_process_group()
