"""A/B of the fused edge-scoring backward (plnlp_edge_mlp_bwd_tf32: dZ1 formed in the GEMM loaders, Hadamard product
re-gathered inside the weight-gradient GEMM) against the unfused backward, forward + backward of EdgeScoreLoss at the
ddi shape (table in L2) and the citation2 shape (232 MB compact table)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops, profiling  # noqa: E402
from tools.microbench import timeit  # noqa: E402

for name, N, H, B, k, drop in (("ddi", 4267, 512, 65536, 3, 0.3), ("citation2", 290000, 200, 65536, 3, 0.0),
                               ("collab-like", 235868, 256, 65536, 3, 0.0)):
    g = torch.Generator(device="cuda").manual_seed(1)
    h0 = torch.randn(N, H, device="cuda", generator=g)
    pos = torch.randint(0, N, (B, 2), device="cuda", generator=g)
    neg = torch.randint(0, N, (B * k, 2), device="cuda", generator=g)
    params0 = [torch.randn(H, H, device="cuda", generator=g) / H ** 0.5, torch.randn(H, device="cuda", generator=g),
               torch.randn(1, H, device="cuda", generator=g) / H ** 0.5, torch.randn(1, device="cuda", generator=g)]
    res = {}
    for mode in ("0", "1"):
        _ops.FUSED_EDGE_BWD = mode
        h = h0.clone().requires_grad_(True)
        params = [p.clone().requires_grad_(True) for p in params0]

        def step():
            h.grad = None
            for p in params:
                p.grad = None
            loss = _ops.edge_score_loss(h, pos, neg, k, "AUC", head="MLP", params=params, drop_p=drop, seed=3)
            loss.backward()
            return loss

        ms = timeit(step, warm=3, iters=10)
        step()
        res[mode] = [h.grad.clone()] + [p.grad.clone() for p in params]
        profiling.enable()
        step()
        prof = profiling.disable()
        print(f"{name}: P={B * (1 + k)} H={H} fused_bwd={mode}: {ms:.3f} ms fwd+bwd", flush=True)
        for kname, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:9]:
            print(f"      {kname:60s} {d['n']:3d} {d['ms']:8.3f} ms")
    worst = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(res["1"], res["0"]))
    print(f"{name}: worst gradient difference fused vs unfused (relative to the tensor's max): {worst:.2e}", flush=True)
_ops.FUSED_EDGE_BWD = "auto"
