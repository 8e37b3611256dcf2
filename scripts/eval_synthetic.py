"""test_co3d.py:93-154 shaped evaluation loop on synthetic data (no dataset or checkpoint exists offline).

For each of `--batches` batches of `--pairs` synthetic image pairs: backbone + lifting (PyTorch), then
the fused hypothesis-and-verification step, geodesic error and Acc@15/30 on the device — one host
sync per batch instead of the reference's per-pair `.item()` (test_co3d.py:149-152).
"""
import argparse, importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modules.model_co3d import Estimator  # noqa: E402
from modules._estimator_base import geodesic_deg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=32)
ap.add_argument("--batches", type=int, default=4)
ap.add_argument("--hyps", type=int, default=50000)   # test_co3d.py:212
args = ap.parse_args()
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)
torch.manual_seed(0)                                  # test_co3d.py:24
model = Estimator({"DATA": {"NUM_ROTA": args.hyps, "BG": True, "SIZE_THR": 0, "ACC_THR": 15}}).to(dev).eval()
proposals = ahv.so3.random_rotations(args.hyps, device=dev)   # once per "category" (test_co3d.py:106)
errs, t_backbone, t_verify = [], 0.0, 0.0
with torch.no_grad():
    for _ in range(args.batches):
        img1, img2 = torch.randn(args.pairs, 3, 256, 256, device=dev), torch.randn(args.pairs, 3, 256, 256, device=dev)
        gt = ahv.so3.sample_rotations(args.pairs, seed=len(errs), device=dev)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        vs, vt = model(img1, img2)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        _, _, R_best, _ = model.predict_rotation(vs, vt, proposals)
        err = geodesic_deg(R_best[:, 0], gt)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        errs.append(err); t_backbone += t1 - t0; t_verify += t2 - t1
err = torch.cat(errs)
print(f"pairs {err.numel()}  mean err {err.mean():.1f} deg  Acc@15 {100 * (err <= 15).float().mean():.1f}%  "
      f"Acc@30 {100 * (err <= 30).float().mean():.1f}%  (random weights: chance level)")
print(f"backbone+lifting {1e3 * t_backbone / args.batches:.1f} ms/batch, hypothesis-and-verification "
      f"{1e3 * t_verify / args.batches:.1f} ms/batch ({args.pairs * args.hyps * args.batches / t_verify:.3g} hyp*pairs/s)")
