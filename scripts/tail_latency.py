"""B=1 latency of everything downstream of the 2D backbone: forward_2d3d (2D embedding,
bidirectional transformer, 3D ResNet block) + fused verification, as ONE CUDA-graph replay (`GraphedTail`),
against the same steps launched eagerly and against the tail with the 3D block left to cuDNN."""
import importlib, json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modules.modules import Feature_Aligner, ResNetBlock_3D, _ResNetBlock  # noqa: E402
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
fa = Feature_Aligner(768, 256, 32, 4, 4).to(dev).eval()
out = {"what": "B=1 post-backbone latency: forward_2d3d + verification (random-init Feature_Aligner 768/256/32, depth 4)"}


def p50(fn, n=100):
    for _ in range(10):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3


for N in (3000, 50000):
    a, b = torch.randn(1, 768, 8, 8, device=dev), torch.randn(1, 768, 8, 8, device=dev)
    R = ahv.so3.sample_rotations(N, seed=1, device=dev)
    gt = ahv.GraphedTail(fa, 1, N, k=1, device=dev)
    gt(a, b, R)
    out[f"N={N}"] = {"graph_replay_p50_us": p50(lambda: gt())}
    ver = ahv.HypothesisVerifier.from_feature_aligner(fa)

    def eager():
        with torch.no_grad():
            vs, vt = fa.forward_2d3d(a, b, random_mask=False, mask_ratio=0.0)
            return ver.score(vs, vt, R, k=1, return_scores=False)
    out[f"N={N}"]["eager_p50_us"] = p50(eager)
    gv = ahv.GraphedVerifier(ver, 1, N, k=1, device=dev)
    with torch.no_grad():
        vs, vt = fa.forward_2d3d(a, b, random_mask=False, mask_ratio=0.0)
    gv(vs, vt, R)
    out[f"N={N}"]["verification_only_graph_p50_us"] = p50(lambda: gv())
# the 3D block alone: one cluster launch against cuDNN's five launches (2 volumes = one pair)
x = torch.randn(2, 32, 8, 8, 8, device=dev)
blk = fa.feature_embedding_3d
with torch.no_grad():
    out["resblock3d_kernel_p50_us"] = p50(lambda: blk(x))
    out["resblock3d_cudnn_p50_us"] = p50(lambda: _ResNetBlock.forward(blk, x))
    x64 = torch.randn(64, 32, 8, 8, 8, device=dev)
    out["resblock3d_kernel_64vol_p50_us"] = p50(lambda: blk(x64))
    out["resblock3d_cudnn_64vol_p50_us"] = p50(lambda: _ResNetBlock.forward(blk, x64))
print(json.dumps(out), flush=True)
