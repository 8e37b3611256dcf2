"""test_co3d.py:93-198 shaped evaluation on synthetic data (no dataset or checkpoint exists offline): per-category
proposals, NP2 frame pairs per sequence, arg-max hypothesis, geodesic error, Acc@15/30 - through
`3dahv_b200.evaluate.evaluate_pairwise`, i.e. the reference's loop with the hypothesis-and-verification idiom replaced
by one fused call per batch of pairs and one host read per category."""
import argparse, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modules.model_co3d import Estimator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sequences", type=int, default=16)
ap.add_argument("--categories", type=int, default=2)
ap.add_argument("--hyps", type=int, default=50000)   # test_co3d.py:212
args = ap.parse_args()
ahv = importlib.import_module("3dahv_b200")
dev = torch.device("cuda", 0)


class SyntheticCo3d:
    """Duck-typed like data_loader_co3d.Co3dDataset as test_co3d.py uses it (:109-113)."""

    def __init__(self, n_seq, seed, n_frames=4):
        g = torch.Generator().manual_seed(seed)
        self.images = [torch.randn(n_frames, 3, 256, 256, generator=g) for _ in range(n_seq)]
        self.R = [torch.linalg.qr(torch.randn(n_frames, 3, 3, generator=g))[0] for _ in range(n_seq)]
        self.n = n_frames

    def __iter__(self):
        for i in range(len(self.images)):
            yield {"n": self.n, "model_id": f"seq{i}"}

    def get_data(self, sequence_name, ids):
        i = int(sequence_name[3:])
        ids = torch.as_tensor(np.asarray(ids))
        return {"image": self.images[i][ids], "R": self.R[i][ids]}


torch.manual_seed(0); np.random.seed(0)               # test_co3d.py:24-25
cfg = {"DATA": {"NUM_ROTA": args.hyps, "BG": True, "SIZE_THR": 0, "ACC_THR": 15}}
model = Estimator(cfg).to(dev).eval()
cats = [f"cat{i}" for i in range(args.categories)]
t0 = time.perf_counter()
errors, acc30, acc15 = ahv.evaluate.evaluate_pairwise(cfg=cfg, model=model, categories=cats, num_frames=2, device=dev, batch_pairs=32,
                                                      get_dataset=lambda cfg, category, split, dataset: SyntheticCo3d(args.sequences, hash(category) % 1000))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pairs = args.categories * args.sequences * 2
print(f"{pairs} pairs x {args.hyps} hypotheses in {dt:.2f} s ({pairs * args.hyps / dt:.3g} hyp*pairs/s including the backbone); "
      f"random weights: chance-level accuracy (mean err {errors['mean']:.1f} deg, Acc@30 {acc30['mean']:.1f} %)")
