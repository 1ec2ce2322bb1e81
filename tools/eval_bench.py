"""Time BaseModel.test (model.py:184-226) at the BASELINE eval shapes, scores and ranking on the
GPU: citation2-shape = 86 596 sources x 1 000 negatives per split, MRR; ddi-shape = 133 489 positives + 100 000
negatives per split, Hits@20/50/100.  Usage: python tools/eval_bench.py [citation2|ddi]   (run under gpurun)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plnlp_b200.graph import CSRGraph  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "citation2"
dev = torch.device("cuda")
cfg = dict(bench.WORKLOADS[which])
torch.manual_seed(0)
data, split = bench.build_workload(cfg, dev, CSRGraph, gcn_normalization)
model = bench.make_model(cfg, dev)
N = cfg["N"]
g = torch.Generator(device=dev).manual_seed(3)
if which == "citation2":
    S, K = 86596, 1000
    for name in ("valid", "test"):
        split[name] = {"source_node": torch.randint(0, N, (S,), generator=g, device=dev),
                       "target_node": torch.randint(0, N, (S,), generator=g, device=dev),
                       "target_node_neg": torch.randint(0, N, (S, K), generator=g, device=dev)}
    metric, pairs = "mrr", 2 * (S + S * K)
else:
    P, Q = 133489, 100000
    for name in ("valid", "test"):
        split[name] = {"edge": torch.randint(0, N, (P, 2), generator=g, device=dev),
                       "edge_neg": torch.randint(0, N, (Q, 2), generator=g, device=dev)}
    metric, pairs = "hits", 2 * (P + Q)
res = model.test(data, split, batch_size=64 * 1024, evaluator=None, eval_metric=metric)     # warm-up
torch.cuda.synchronize()
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = model.test(data, split, batch_size=64 * 1024, evaluator=None, eval_metric=metric)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"workload": cfg["name"], "call": "BaseModel.test (encode once, score valid + test, rank on the GPU)",
                  "pairs_scored": pairs, "ms": round(ms, 2), "wall_s": round(time.time() - t0, 3),
                  "pairs_per_s": round(pairs / ms * 1e3), "result": {k: [round(v, 5) for v in vs] for k, vs in res.items()}}))
