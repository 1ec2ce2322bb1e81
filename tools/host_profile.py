"""Where does the HOST time of a training step go?  At 8 GPUs the per-rank kernels of the citation2-shape step add up
to ~5 ms, and python + launch calls + host reads decide the step time.  This runs the single-GPU step on a graph
scaled down by SCALE (so the GPU is as lightly loaded as one of 8 ranks) under cProfile.
Usage: python tools/host_profile.py [SCALE=8] [steps=30]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from plnlp_b200.graph import CSRGraph  # noqa: E402
from plnlp_b200.utils import gcn_normalization, get_pos_neg_edges  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = dict(bench.WORKLOADS["citation2"])
cfg["N"] //= scale
cfg["E"] //= scale
dev = torch.device("cuda")
torch.manual_seed(0)
data, split = bench.build_workload(cfg, dev, CSRGraph, gcn_normalization)
model = bench.make_model(cfg, dev)
model.num_nodes = cfg["N"] * scale            # a batch still counts as touching a small part of the node set
B, k = cfg["batch"] // scale, cfg["num_neg"]
pos_all = torch.stack([split["train"]["source_node"], split["train"]["target_node"]], 1)


def run(n):
    idx = torch.randint(0, pos_all.size(0), (n * B,), device=dev)
    pos, neg = get_pos_neg_edges("train", {"train": {"edge": pos_all[idx]}}, edge_index=data.edge_index,
                                 num_nodes=cfg["N"], neg_sampler_name="local", num_neg=k, device=dev)
    model.encoder.train(); model.predictor.train()
    model.run_batches(data, ((pos[i * B:(i + 1) * B], neg[i * B:(i + 1) * B].reshape(-1, 2), None) for i in range(n)), k)


run(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
run(steps)
host = (time.perf_counter() - t0) * 1e3 / steps
e1.record()
torch.cuda.synchronize()
print(f"scale 1/{scale}: device {e0.elapsed_time(e1) / steps:.3f} ms/step, host enqueue {host:.3f} ms/step", flush=True)
pr = cProfile.Profile()
pr.enable()
run(steps)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
