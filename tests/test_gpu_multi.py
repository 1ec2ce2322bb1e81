"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped otherwise): the row-partitioned SpMM with the CUDA
kernel as the local operator over NCCL, and one full training step of the partitioned / data-parallel
model, against the single-GPU result (index work identical, fp32 within
tolerance because the summation order changes)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import rand_graph, rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Data:
    pass


def _model(N, feats, dev, enc):
    from plnlp_b200.model import BaseModel
    torch.manual_seed(7)
    m = BaseModel(lr=0.01, dropout=0.0, grad_clip_norm=-1.0, gnn_num_layers=2, mlp_num_layers=2,
                  emb_hidden_channels=16, gnn_hidden_channels=32, mlp_hidden_channels=32, num_nodes=N,
                  num_node_feats=feats, gnn_encoder_name=enc, predictor_name="MLP", loss_func="AUC",
                  optimizer_name="SGD", device=dev, use_node_feats=feats > 0, train_node_emb=True)
    m.param_init()
    return m


def _worker(rank, ws, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    try:
        from plnlp_b200 import _ops, parallel
        from plnlp_b200.graph import CSRGraph
        from plnlp_b200.utils import gcn_normalization
        N, F = 301, 24
        ei, _ = rand_graph(N, 3000, seed=9, hub=True)
        ei = ei[:, ei[0] != ei[1]]
        adj = gcn_normalization(CSRGraph.from_edge_index(torch.cat([ei, ei.flip(0)], 1).to(dev), None, N).to_symmetric())
        g = torch.Generator().manual_seed(1)
        x, gout = torch.randn(N, F, generator=g).to(dev), torch.randn(N, F, generator=g).to(dev)
        lo, hi = parallel.row_block(N, rank, ws)
        blk = parallel.block_size(N, ws)
        sadj = parallel.shard_graph(adj, rank, ws, CSRGraph)
        # ---- partitioned SpMM vs single GPU
        xr = x.clone().requires_grad_(True)
        y = _ops.spmm(adj, xr, "sum")
        y.backward(gout)
        xl = x[lo:hi].clone().requires_grad_(True)
        yl = _ops.spmm(sadj, xl, "sum")
        gl = torch.zeros(blk, F, device=dev)
        gl[: hi - lo] = gout[lo:hi]
        yl.backward(gl)
        e_fwd, e_bwd = rel_err(yl[: hi - lo], y[lo:hi]), rel_err(xl.grad, xr.grad[lo:hi])
        # ---- one training step: partitioned model vs single-GPU model on the union of the ranks' batches
        B, k = 64, 2
        gg = torch.Generator().manual_seed(2)
        pos_all = torch.randint(0, N, (ws * B, 2), generator=gg).to(dev)
        neg_all = torch.randint(0, N, (ws * B, k, 2), generator=gg).to(dev)
        feats = torch.randn(N, 8, generator=gg).to(dev)
        # fresh adjacency objects with the dense tensor-core path off, so that the CSR kernels and the
        # aggregate-buffer path of GCNConv (single-device and row-partitioned) are what is compared
        from plnlp_b200 import graph
        graph.DENSE_SPMM = False
        adj = CSRGraph(*adj.csr(), (N, N))
        sadj = parallel.shard_graph(adj, rank, ws, CSRGraph)
        single = _model(N, 8, dev, "GCN")
        d1 = Data(); d1.adj_t, d1.x, d1.edge_index = adj, feats, None
        single.encoder.train(); single.predictor.train()
        single.optimizer = torch.optim.SGD(single.para_list, lr=0.0)
        loss1 = single.train_batch(d1, pos_all, neg_all.reshape(-1, 2), k)
        part = _model(blk, 8, dev, "GCN")
        with torch.no_grad():          # same parameters: replicated weights, row block of the embedding
            for a, b in zip(part.encoder.parameters(), single.encoder.parameters()):
                a.copy_(b)
            for a, b in zip(part.predictor.parameters(), single.predictor.parameters()):
                a.copy_(b)
            part.emb.weight.zero_()
            part.emb.weight[: hi - lo].copy_(single.emb.weight[lo:hi])
        part.world_size, part.rank, part.partitioned = ws, rank, True
        part.optimizer = torch.optim.SGD(part.para_list, lr=0.0)
        d2 = Data(); d2.adj_t, d2.x, d2.edge_index = sadj, parallel.pad_rows(feats[lo:hi].contiguous(), blk), None
        part.encoder.train(); part.predictor.train()
        errs = {}
        for mode in ("rows", "allgather"):      # compact endpoint-row exchange (default) and whole-matrix all-gather
            parallel.EXCHANGE = mode
            loss2 = part.train_batch(d2, pos_all[rank * B:(rank + 1) * B],
                                     neg_all[rank * B:(rank + 1) * B].reshape(-1, 2), k)
            tot = loss2.clone()
            dist.all_reduce(tot)
            errs[mode + ".loss"] = abs(float(tot) - float(loss1)) / abs(float(loss1))
            errs[mode + ".emb"] = rel_err(part.emb.weight.grad[: hi - lo], single.emb.weight.grad[lo:hi])
            for (n1, a), b in zip(part.encoder.named_parameters(), single.encoder.parameters()):
                errs[mode + ".enc." + n1] = rel_err(a.grad, b.grad)
            for (n1, a), b in zip(part.predictor.named_parameters(), single.predictor.parameters()):
                errs[mode + ".pred." + n1] = rel_err(a.grad, b.grad)
        parallel.EXCHANGE = "rows"
        assert "_plnlp_agg_buffer" in sadj.__dict__ and "_plnlp_agg_buffer" in adj.__dict__
        ret[rank] = {"spmm_fwd": e_fwd, "spmm_bwd": e_bwd, **errs}
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_partitioned_matches_single_gpu_nccl_ws2():
    ws = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(ws, _free_port(), ret), nprocs=ws, join=True)
    for r in range(ws):
        for k, v in ret[r].items():
            # predictor gradients are sums over pairs weighted by d loss/d score, which sums to exactly 0
            # for the AUC loss: heavy cancellation, and the two runs add the pairs in different orders
            tol = 2e-3 if ".pred." in k else 2e-5
            assert v < tol, (r, k, v)


def _restricted_worker(rank, ws, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    try:
        from plnlp_b200 import graph, parallel
        from plnlp_b200.graph import CSRGraph
        from plnlp_b200.utils import gcn_normalization
        graph.DENSE_SPMM = False
        N = 301
        ei, _ = rand_graph(N, 3000, seed=9, hub=True)
        ei = ei[:, ei[0] != ei[1]]
        adj = gcn_normalization(CSRGraph.from_edge_index(torch.cat([ei, ei.flip(0)], 1).to(dev), None, N).to_symmetric())
        lo, hi = parallel.row_block(N, rank, ws)
        blk = parallel.block_size(N, ws)
        sadj = parallel.shard_graph(adj, rank, ws, CSRGraph)
        B, k = 16, 2                                    # few pairs: the batches touch a small part of the nodes
        gg = torch.Generator().manual_seed(2)
        pos_all = torch.randint(0, N, (ws * B, 2), generator=gg).to(dev)
        neg_all = torch.randint(0, N, (ws * B, k, 2), generator=gg).to(dev)
        feats = torch.randn(N, 8, generator=gg).to(dev)
        single = _model(N, 8, dev, "GCN")
        d1 = Data(); d1.adj_t, d1.x, d1.edge_index = adj, feats, None
        single.encoder.train(); single.predictor.train()
        single.optimizer = torch.optim.SGD(single.para_list, lr=0.0)
        loss1 = single.train_batch(d1, pos_all, neg_all.reshape(-1, 2), k)
        part = _model(blk, 8, dev, "GCN")
        with torch.no_grad():
            for a, b in zip(part.encoder.parameters(), single.encoder.parameters()):
                a.copy_(b)
            for a, b in zip(part.predictor.parameters(), single.predictor.parameters()):
                a.copy_(b)
            part.emb.weight.zero_()
            part.emb.weight[: hi - lo].copy_(single.emb.weight[lo:hi])
        part.world_size, part.rank, part.partitioned = ws, rank, True
        part.num_nodes = 10 ** 9                         # "a batch touches a small part of the node set"
        part.optimizer = torch.optim.SGD(part.para_list, lr=0.0)
        d2 = Data(); d2.adj_t, d2.x, d2.edge_index = sadj, parallel.pad_rows(feats[lo:hi].contiguous(), blk), None
        part.encoder.train(); part.predictor.train()
        parallel.RESTRICT_LAST, parallel.EXCHANGE = True, "rows"
        loss2 = part.train_batch(d2, pos_all[rank * B:(rank + 1) * B], neg_all[rank * B:(rank + 1) * B].reshape(-1, 2), k)
        tot = loss2.clone()
        dist.all_reduce(tot)
        errs = {"loss": abs(float(tot) - float(loss1)) / abs(float(loss1)),
                "emb": rel_err(part.emb.weight.grad[: hi - lo], single.emb.weight.grad[lo:hi])}
        for (n1, a), b in zip(part.encoder.named_parameters(), single.encoder.parameters()):
            errs["enc." + n1] = rel_err(a.grad, b.grad)
        for (n1, a), b in zip(part.predictor.named_parameters(), single.predictor.parameters()):
            errs["pred." + n1] = rel_err(a.grad, b.grad)
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.skipif(os.environ.get("PLNLP_RUN_STAGED") != "1",
                    reason="PLNLP_PARTITIONED_RESTRICT (off by default) was written after the round's GPU minutes were "
                           "spent: bookkeeping verified with gloo on CPU only.  Run with PLNLP_RUN_STAGED=1 on 2 GPUs "
                           "(under a timeout: a mismatch in the collective sequence would hang), then drop this mark")
def test_partitioned_restricted_last_layer_nccl_ws2():
    """requests first, then every owner computes the last conv only for the requested rows
    (parallel.exchange_row_requests / pspmm_rows / serve_rows) vs the single-GPU step"""
    ws = 2
    ret = mp.Manager().dict()
    mp.spawn(_restricted_worker, args=(ws, _free_port(), ret), nprocs=ws, join=True)
    for r in range(ws):
        for k, v in ret[r].items():
            tol = 2e-3 if k.startswith("pred.") else 2e-5
            assert v < tol, (r, k, v)
