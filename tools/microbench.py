"""Per-kernel timings at the BASELINE shapes (CUDA events on the launching stream).
Usage: python tools/microbench.py [ddi|citation2|sweep] ...   (run under gpurun)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, structure_of  # noqa: E402

DEV = torch.device("cuda")
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = PEAKS.get("hbm_gbs", 6650.0)


def timeit(fn, warm=2, iters=10, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def undirected_graph(N, E, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lo = torch.randint(0, N, (int(E * 1.25),), generator=g, device=DEV)
    hi = torch.randint(0, N, (int(E * 1.25),), generator=g, device=DEV)
    key = torch.unique(torch.minimum(lo, hi) * N + torch.maximum(lo, hi))
    key = key[(key // N) != (key % N)]
    key = key[torch.randperm(key.numel(), device=DEV, generator=g)[:E]]
    ei = torch.stack([key // N, key % N])
    return torch.cat([ei, ei.flip(0)], 1)


def powerlaw_graph(N, E, seed, alpha=2.1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    # in-degree power law: destination drawn from a Zipf-like distribution over a random node order
    u = torch.rand(E, generator=g, device=DEV)
    dst_rank = (N * u.pow(alpha)).long().clamp(max=N - 1)
    perm = torch.randperm(N, device=DEV, generator=g)
    dst = perm[dst_rank]
    src = torch.randint(0, N, (E,), generator=g, device=DEV)
    return torch.stack([src, dst])


def bench_spmm(name, adj, F, reduce, flush):
    st = structure_of(adj)
    plan = st.fwd_noval if reduce == "mean" else st.fwd
    x = torch.randn(adj.size(1), F, device=DEV)
    ms = timeit(lambda: _ops.spmm_raw(plan, x, use_val=reduce != "mean", div_rows=reduce == "mean"), flush=flush)
    by = plan.alg_bytes(F) if reduce != "mean" else plan.alg_bytes(F) - (plan.nnz * 4 if plan.val is not None else 0)
    print(f"[spmm] {name} F={F} {reduce}: {ms:.3f} ms  alg {by/1e9:.2f} GB -> {by/ms/1e6:.0f} GB/s "
          f"({by/ms/1e6/HBM*100:.0f}% of measured HBM {HBM:.0f}) items={plan.n_items} chunk={plan.chunk} nfix={plan.n_fix}",
          flush=True)
    return ms


def bench_gemm(M, N, K, ta=False, tb=True, **kw):
    A = torch.randn((K, M) if ta else (M, K), device=DEV)
    B = torch.randn((N, K) if tb else (K, N), device=DEV)
    ms = timeit(lambda: _ops.gemm_raw(A, B, transa=ta, transb=tb, **kw))
    fl = 2.0 * M * N * K
    print(f"[gemm {kw.get('backend', 'ffma')}] M={M} N={N} K={K} ta={int(ta)} tb={int(tb)}: {ms:.3f} ms -> {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    return ms


def ddi():
    N, E, H, B, k = 4267, 1067911, 512, 65536, 3
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
    adj = CSRGraph.from_edge_index(undirected_graph(N, E, 0), None, N)
    bench_spmm("ddi", adj, H, "mean", None)
    bench_spmm("ddi(L2 flushed)", adj, H, "mean", flush)
    P = B * (1 + k)
    for be in (["ffma"] if "noffma" not in sys.argv else []) + ["tf32x3", "tf32x3c2", "tf32", "tf32c2"]:
        bench_gemm(N, H, H, backend=be)
        bench_gemm(P, H, H, backend=be)                       # predictor layer 1 fwd
        bench_gemm(P, H, H, tb=False, backend=be)             # dA0 = dZ1 @ W1
        bench_gemm(H, H, P, ta=True, tb=False, backend=be)    # dW1 = dZ1^T @ A0 (split-k)
    h = torch.randn(N, H, device=DEV)
    edges = torch.randint(0, N, (P, 2), device=DEV)
    ms = timeit(lambda: _ops.gather_hadamard_raw(h, edges))
    print(f"[gather_hadamard] P={P} H={H}: {ms:.3f} ms -> {(P*H*4*3)/ms/1e6:.0f} GB/s (2 gathered rows + 1 written)")
    da = torch.randn(P, H, device=DEV)
    for mode in ("sorted", "atomic"):
        ms = timeit(lambda: _ops.edge_scatter_raw(h, edges, da=da, mode=mode))
        print(f"[edge_scatter {mode}] {ms:.3f} ms")
    a = torch.relu(torch.randn(P, H, device=DEV))
    w, b, ds = torch.randn(1, H, device=DEV), torch.randn(1, device=DEV), torch.randn(P, device=DEV)
    ms = timeit(lambda: _ops.mlp_out_fwd_raw(a, w, b))
    print(f"[mlp_out_fwd] {ms:.3f} ms -> {P*H*4/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: _ops.mlp_out_bwd_raw(a, w, ds, True, 1.0))
    print(f"[mlp_out_bwd] {ms:.3f} ms -> {P*H*8/ms/1e6:.0f} GB/s")
    pos, neg = torch.randn(B, device=DEV), torch.randn(B * k, device=DEV)
    ms = timeit(lambda: _ops.pair_loss_raw(0, pos, neg, k))
    print(f"[pair_loss] {ms:.3f} ms")


def citation2(scale=1.0):
    N, E, F = int(2927963 * scale), int(30561187 * scale), 200
    ei = powerlaw_graph(N, E, 1)
    t0 = time.time()
    adj = CSRGraph.from_edge_index(ei, None, N).to_symmetric()
    from plnlp_b200.utils import gcn_normalization
    adj = gcn_normalization(adj)
    st = structure_of(adj)
    torch.cuda.synchronize()
    print(f"citation2-shape graph: N={N} nnz={adj.nnz()} built in {time.time()-t0:.1f}s symmetric={st.symmetric} "
          f"max_deg={int((adj.csr()[0][1:]-adj.csr()[0][:-1]).max())}")
    bench_spmm("citation2", adj, F, "sum", None)
    bench_spmm("citation2", adj, 128, "sum", None)
    bench_spmm("citation2", adj, 256, "sum", None)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ddi"]
    print("device:", torch.cuda.get_device_name(0), "HBM peak (measured):", HBM)
    if "ddi" in which:
        ddi()
    if "citation2" in which:
        citation2()
    if "citation2_small" in which:
        citation2(0.25)


def gemm_epilogues():
    """isolate the epilogue cost of the predictor forward GEMM"""
    P, H = 262144, 512
    A, W, b = torch.randn(P, H, device=DEV), torch.randn(H, H, device=DEV), torch.randn(H, device=DEV)
    for name, kw in (("plain", {}), ("bias", dict(bias=b)), ("bias+relu", dict(bias=b, act=1)),
                     ("bias+relu+drop", dict(bias=b, act=1, drop_p=0.3, seed=7))):
        ms = timeit(lambda: _ops.gemm_raw(A, W, transb=True, backend="tf32x3", **kw))
        print(f"[gemm tf32x3 epilogue {name}] {ms:.3f} ms -> {2.0*P*H*H/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__" and "epi" in sys.argv:
    gemm_epilogues()


def collab_dot():
    """BASELINE config 3 shape: DOT head over a 241 MB h (does not fit L2): fused gather+dot forward and
    the sorted-scatter backward against the gather roofline (fwd 2*H*4+12 B/pair, bwd +4*H*4)."""
    N, H, P = 235868, 256, 131072
    h = torch.randn(N, H, device=DEV)
    edges = torch.randint(0, N, (P, 2), device=DEV)
    ds = torch.randn(P, device=DEV)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
    ms = timeit(lambda: _ops.edge_dot_raw(h, edges), flush=flush)
    by = P * (2 * H * 4 + 12)
    print(f"[edge_dot fwd collab-shape] {ms:.4f} ms -> {by/ms/1e6:.0f} GB/s ({by/ms/1e6/HBM*100:.0f}% of HBM)")
    for mode in ("sorted", "atomic"):
        ms = timeit(lambda: _ops.edge_scatter_raw(h, edges, dscore=ds, mode=mode), flush=flush)
        by = P * (4 * H * 4 + 20) + N * H * 4
        print(f"[edge_dot bwd {mode} collab-shape] {ms:.4f} ms -> {by/ms/1e6:.0f} GB/s ({by/ms/1e6/HBM*100:.0f}% of HBM; "
              f"includes the N*H*4 dense grad_h write)")


def sweep():
    """BASELINE config 5: directed power-law graphs 10 M - 200 M edges, source matrix >= 4x L2, hidden 64-512,
    fp32 and bf16 storage: SpMM forward and backward (the transposed plan), edge-score DOT forward + backward,
    edge-score MLP forward (fused tcgen05 kernel), each against the HBM roofline of SURVEY 8d's byte model."""
    P = 1 << 20
    emax = max([int(a[5:]) for a in sys.argv if a.startswith("emax=")] or [200]) * 1_000_000
    for E, N in ((10_000_000, 2_000_000), (50_000_000, 4_000_000), (100_000_000, 4_000_000), (200_000_000, 4_000_000)):
        if E > emax:
            continue
        ei = powerlaw_graph(N, E, 5)
        adj = CSRGraph.from_edge_index(ei, None, N)
        del ei
        st = structure_of(adj)
        torch.cuda.synchronize()
        tag = f"powerlaw E={E // 1_000_000}M N={N // 1_000_000}M"
        for F in (64, 128, 256, 512):
            x32 = torch.randn(N, F, device=DEV)
            for dt in (torch.float32, torch.bfloat16):
                x = x32.to(dt)
                es = x.element_size()
                for name, plan in (("fwd", st.fwd), ("bwd", st.bwd)):
                    ms = timeit(lambda: _ops.spmm_raw(plan, x, use_val=False, div_rows=False), iters=3)
                    by = plan.alg_bytes(F, es)
                    print(f"[spmm {name}] {tag} F={F} {str(dt)[6:]}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s "
                          f"({by / ms / 1e6 / HBM * 100:.0f}% of HBM)", flush=True)
                del x
            if E == 50_000_000:         # edge scoring does not depend on the graph: once per width
                edges = torch.randint(0, N, (P, 2), device=DEV)
                ds = torch.randn(P, device=DEV)
                ms = timeit(lambda: _ops.edge_dot_raw(x32, edges), iters=5)
                by = P * (2 * F * 4 + 12)
                print(f"[edge_dot fwd] N={N // 1_000_000}M P={P} H={F}: {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s "
                      f"({by / ms / 1e6 / HBM * 100:.0f}% of HBM)", flush=True)
                ms = timeit(lambda: _ops.edge_scatter_raw(x32, edges, dscore=ds, mode="atomic"), iters=5)
                by = P * (4 * F * 4 + 20) + N * F * 4
                print(f"[edge_dot bwd atomic] H={F}: {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s "
                      f"({by / ms / 1e6 / HBM * 100:.0f}% of HBM, incl. the dense grad_h zero-fill + write)", flush=True)
                W1, b1 = torch.randn(F, F, device=DEV) / F ** 0.5, torch.randn(F, device=DEV)
                w2, b2 = torch.randn(F, device=DEV), torch.randn(1, device=DEV)
                ms = timeit(lambda: _ops.edge_mlp_fwd_raw(x32, edges, W1, b1, w2, b2, need_a1=False), iters=5)
                by, fl = P * (2 * F * 4 + 12), 2.0 * P * F * F + 2.0 * P * F
                print(f"[edge_mlp fwd fused] H={F}: {ms:.4f} ms  gather {by / ms / 1e6:.0f} GB/s "
                      f"({by / ms / 1e6 / HBM * 100:.0f}% of HBM)  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
            del x32
        del adj, st
        torch.cuda.empty_cache()


if __name__ == "__main__" and "collab" in sys.argv:
    collab_dot()
if __name__ == "__main__" and "sweep" in sys.argv:
    sweep()
