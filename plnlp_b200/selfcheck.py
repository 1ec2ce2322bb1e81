"""Multi-GPU self check: one optimisation step of the ROW-PARTITIONED model against the single-GPU model on the
same parameters, graph and edge batches (the reference has no multi-device code, so the parity
target of the partitioned run is the 1-GPU result -- index work identical, fp32 within tolerance because the
summation order changes).

Used by ``bench.py`` (untimed, before the timed region of every N > 1 run, so the scaling record carries parity
evidence from the very ranks that were timed) and by ``tests/test_gpu_multi.py`` (which adds an fp64 oracle as the
yardstick for the ill-conditioned gradients).  Everything here runs on the CUDA kernels; nothing is imported from
``oracle/``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import parallel
from .graph import CSRGraph
from .utils import gcn_normalization


class _Data:
    pass


def synthetic_graph(n, e, device, seed=9):
    """directed power-law-ish edge list -> symmetrised, unit diagonal, D^-1/2 (A + I) D^-1/2 (main.py:109-110,
    177-179); the last nodes stay isolated so that padding / empty rows are exercised"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    hi = max(n - 3, 1)
    dst = (hi * torch.rand(e, generator=g).pow(2.0)).long().clamp(max=hi - 1)
    src = torch.randint(0, hi, (e,), generator=g)
    keep = src != dst
    ei = torch.stack([src[keep], dst[keep]]).to(device)
    adj = CSRGraph.from_edge_index(ei, None, n).to_symmetric()
    return gcn_normalization(adj)


def _model(n_rows, feats, emb, hid, device, encoder, loss):
    from .model import BaseModel
    m = BaseModel(lr=0.01, dropout=0.0, grad_clip_norm=-1.0, gnn_num_layers=2, mlp_num_layers=2,
                  emb_hidden_channels=emb, gnn_hidden_channels=hid, mlp_hidden_channels=hid, num_nodes=n_rows,
                  num_node_feats=feats, gnn_encoder_name=encoder, predictor_name="MLP", loss_func=loss,
                  optimizer_name="SGD", device=device, use_node_feats=feats > 0, train_node_emb=True)
    m.param_init()
    m.optimizer = torch.optim.SGD(m.para_list, lr=0.0)       # the step leaves the parameters where they were
    m.encoder.train()
    m.predictor.train()
    return m


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def partitioned_step(rank, ws, device, adj=None, n=60000, e=600000, feats=16, emb=8, hid=32, batch=256, k=3,
                     encoder="GCN", loss="AUC", force_sparse=None, seed=7, keep=False):
    """-> dict of relative errors (loss, every gradient tensor) of the partitioned step vs the single-GPU step.
    ``batch`` positives per rank; ``force_sparse`` True / False overrides whether the batch counts as touching a
    small part of the node set (the row-restricted last conv) -- None keeps the model's own rule.  ``keep``:
    also return the gradients ('single' / 'part' dicts of CPU tensors, emb rows of this rank's block)."""
    from . import graph
    dense_was = graph.DENSE_SPMM
    graph.DENSE_SPMM = False                                 # the CSR kernels are what a partitioned run uses
    try:
        if adj is None:
            adj = synthetic_graph(n, e, device)
        n = adj.size(0)
        blk = parallel.block_size(n, ws)
        lo, hi = parallel.row_block(n, rank, ws)
        sadj = parallel.shard_graph(adj, rank, ws, CSRGraph)
        g = torch.Generator().manual_seed(seed)
        pos_all = torch.randint(0, n, (ws * batch, 2), generator=g).to(device)
        neg_all = torch.randint(0, n, (ws * batch, k, 2), generator=g).to(device)
        x = torch.randn(n, feats, generator=g).to(device) if feats else None
        torch.manual_seed(seed)
        single = _model(n, feats, emb, hid, device, encoder, loss)
        part = _model(blk, feats, emb, hid, device, encoder, loss)
        with torch.no_grad():          # same parameters: replicated weights, row block of the embedding
            for a, b in zip(part.encoder.parameters(), single.encoder.parameters()):
                a.copy_(b)
            for a, b in zip(part.predictor.parameters(), single.predictor.parameters()):
                a.copy_(b)
            part.emb.weight.zero_()
            part.emb.weight[: hi - lo].copy_(single.emb.weight[lo:hi])
        part.world_size, part.rank, part.partitioned = ws, rank, True
        part.num_nodes = n
        if force_sparse is not None:
            single.num_nodes = part.num_nodes = (10 ** 12 if force_sparse else 1)
        d1, d2 = _Data(), _Data()
        d1.adj_t, d1.x, d1.edge_index = adj, x, None
        d2.adj_t, d2.edge_index = sadj, None
        d2.x = None if x is None else parallel.pad_rows(x[lo:hi].contiguous(), blk)
        loss1 = single.train_batch(d1, pos_all, neg_all.reshape(-1, 2), k)
        sl = slice(rank * batch, (rank + 1) * batch)
        loss2 = part.train_batch(d2, pos_all[sl], neg_all[sl].reshape(-1, 2), k)
        tot = loss2.detach().clone().double()
        dist.all_reduce(tot)
        if loss in ("CE", "InfoNCE", "LogRank"):
            tot /= ws
        errs = {"loss": abs(float(tot) - float(loss1)) / max(abs(float(loss1)), 1e-30),
                "emb": _rel(part.emb.weight.grad[: hi - lo], single.emb.weight.grad[lo:hi])}
        grads = {"single": {}, "part": {}}
        pairs = [("enc." + nm, a, b) for (nm, a), b in zip(part.encoder.named_parameters(), single.encoder.parameters())]
        pairs += [("pred." + nm, a, b) for (nm, a), b in zip(part.predictor.named_parameters(), single.predictor.parameters())]
        # absolute error of every gradient tensor relative to the LARGEST gradient magnitude of the model: the
        # yardstick for tensors whose exact value is (nearly) zero by cancellation (bias gradients of a pairwise
        # loss whose d loss / d score sums to zero)
        scale = float(single.emb.weight.grad.abs().max())
        worst = float((part.emb.weight.grad[: hi - lo].double() - single.emb.weight.grad[lo:hi].double()).abs().max())
        for key, a, b in pairs:
            errs[key] = _rel(a.grad, b.grad)
            scale = max(scale, float(b.grad.abs().max()))
            worst = max(worst, float((a.grad.double() - b.grad.double()).abs().max()))
            if keep:
                grads["single"][key], grads["part"][key] = b.grad.cpu(), a.grad.cpu()
        errs["max_abs_over_model_scale"] = worst / max(scale, 1e-30)
        # fp32 noise floor of the comparison: the SAME single-GPU step with the pairs of the batch in another order
        # (mathematically identical loss and gradients; only the summation order inside the scatter / reduction
        # kernels changes, as it does between 1 and N ranks)
        ref = [b.grad.detach().clone() for _, _, b in pairs] + [single.emb.weight.grad.detach().clone()]
        perm = torch.randperm(pos_all.size(0), generator=torch.Generator().manual_seed(seed + 1)).to(device)
        single.train_batch(d1, pos_all[perm], neg_all[perm].reshape(-1, 2), k)
        again = [b.grad for _, _, b in pairs] + [single.emb.weight.grad]
        errs["noise_floor_over_model_scale"] = max(
            float((a.double() - b.double()).abs().max()) for a, b in zip(again, ref)) / max(scale, 1e-30)
        if keep:
            grads["single"]["emb"] = single.emb.weight.grad[lo:hi].cpu()
            grads["part"]["emb"] = part.emb.weight.grad[: hi - lo].cpu()
            grads["single"]["loss"], grads["part"]["loss"] = float(loss1), float(tot)
            grads["inputs"] = {"pos": pos_all.cpu(), "neg": neg_all.cpu(), "x": None if x is None else x.cpu(),
                               "state": {"emb": single.emb.weight.detach().cpu(),
                                         **{"enc." + k_[len("convs."):]: v.detach().cpu()
                                            for k_, v in single.encoder.named_parameters()},
                                         **{"pred." + k_[len("lins."):]: v.detach().cpu()
                                            for k_, v in single.predictor.named_parameters()}}}
            errs["_tensors"] = grads
        return errs
    finally:
        graph.DENSE_SPMM = dense_was


def summarize(errs):
    """the two numbers bench.py prints: loss error and the worst gradient error (on the model-wide scale)"""
    per_tensor = {k: v for k, v in errs.items() if not k.startswith("_")
                  and k not in ("loss", "max_abs_over_model_scale", "noise_floor_over_model_scale")}
    return {"loss_rel": errs["loss"], "max_grad_rel": errs["max_abs_over_model_scale"],
            "fp32_noise_floor": errs["noise_floor_over_model_scale"],
            "max_grad_rel_per_tensor": max(per_tensor.values()),
            "worst_tensor": max(per_tensor, key=per_tensor.get)}
