"""Cross-encoder throughput (BASELINE configs[4]: 1024 queries x 100 candidates, MiniLM-L12) on one GPU,
next to the float32 transformers forward on the host cores for a bounded sample."""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8192)
ap.add_argument("--mean-len", type=int, default=200)
ap.add_argument("--cpu-pairs", type=int, default=64)
ap.add_argument("--tokens-per-call", type=int, default=1 << 18)
args = ap.parse_args()

from oracle import rerank as orr
from raglite_b200._xenc import CrossEncoderEngine

model = orr.seeded_model(seed=0)
eng = CrossEncoderEngine.from_hf(model, max_tokens_per_call=args.tokens_per_call)
rng = np.random.default_rng(0)
lens = np.clip(rng.normal(args.mean_len, 60, size=args.pairs).astype(int), 32, 512)
ids = [rng.integers(1000, 30000, size=L).astype(np.int32) for L in lens]
types = [np.r_[np.zeros(12, np.int32), np.ones(L - 12, np.int32)] for L in lens]
eng.score_tokens(ids[:256], types[:256])
torch.cuda.synchronize()
t0 = time.perf_counter()
logits, scores = eng.score_tokens(ids, types)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
T = int(lens.sum())
H, F, Lyr = 384, 1536, 12
flops = Lyr * (2.0 * T * (3 * H * H + H * H + 2 * H * F) + 4.0 * float((lens.astype(np.float64) ** 2).sum()) * H)
n = args.cpu_pairs
t0 = time.perf_counter()
ref = orr.hf_logits(model, ids[:n], types[:n])
cpu_dt = time.perf_counter() - t0
print(json.dumps({
    "metric": "cross-encoder pairs/sec (MiniLM-L12-H384, packed varlen, fp16 tensor cores)", "pairs": args.pairs,
    "tokens": T, "mean_len": float(lens.mean()), "gpu_pairs_per_s": args.pairs / dt, "gpu_tokens_per_s": T / dt,
    "gpu_tflops": flops / dt / 1e12, "seconds": dt, "c5_seconds_extrapolated": 102400 / (args.pairs / dt),
    "cpu_pairs_per_s": n / cpu_dt, "cpu_threads": torch.get_num_threads(), "cpu_sample_pairs": n,
    "max_abs_logit_err_vs_fp32": float(np.abs(logits[:n] - ref).max()),
}))
