"""CPU tests (-m "not gpu") of plnlp_b200/shims.py: the stand-ins for the third-party calls
of the reference's main.py against the oracle's restatement of the same calls, and the REAL main.py executed
unchanged through them up to the point where it constructs BaseModel (which refuses a CUDA-less box)."""
import os
import runpy
import sys

import pytest
import torch

from oracle import sparse
from plnlp_b200 import shims

MAIN = "/root/reference/main.py"


def test_to_sparse_tensor_and_constructor_match_oracle():
    g = torch.Generator().manual_seed(1)
    N, E = 50, 400
    ei = torch.randint(0, N, (2, E), generator=g)             # duplicates and self loops included
    w = torch.rand(E, generator=g)
    data = shims.ToSparseTensor()(shims.Data(edge_index=ei.clone(), edge_weight=w.clone(), num_nodes=N, x=None))
    assert not hasattr(data, "edge_index") and not hasattr(data, "edge_weight")
    want = sparse.to_sparse_tensor(ei, w, N)
    for a, b in zip(data.adj_t.csr(), want.csr()):
        assert torch.equal(a, b)
    # SparseTensor(row=, col=, value=) as main.py:124-126 builds it
    adj = shims.SparseTensor(row=ei[0], col=ei[1], value=w, sparse_sizes=(N, N))
    want = sparse.SparseTensor(row=ei[0], col=ei[1], value=w, sparse_sizes=(N, N))
    for a, b in zip(adj.csr(), want.csr()):
        assert torch.equal(a, b)
    assert data.num_features == 0 and shims.Data(x=torch.zeros(3, 7)).num_features == 7


def test_to_undirected_and_coalesce_brute_force():
    g = torch.Generator().manual_seed(2)
    N, E = 30, 200
    ei = torch.randint(0, N, (2, E), generator=g)
    w = torch.randint(1, 5, (E,), generator=g).float()
    idx, val = shims.to_undirected(ei, w, reduce="add")
    ref = {}
    for (s, d), x in zip(ei.t().tolist(), w.tolist()):
        for a, b in ((s, d), (d, s)):
            ref[(a, b)] = ref.get((a, b), 0.0) + x
    keys = sorted(ref)
    assert idx.t().tolist() == [list(k) for k in keys]
    assert torch.allclose(val, torch.tensor([ref[k] for k in keys]))
    assert torch.equal(shims.to_undirected(ei), idx)                      # without an attribute: the index only
    cidx, cval = shims.coalesce(ei, w, N, N)
    one = {}
    for (s, d), x in zip(ei.t().tolist(), w.tolist()):
        one[(s, d)] = one.get((s, d), 0.0) + x
    assert cidx.t().tolist() == [list(k) for k in sorted(one)]
    assert torch.allclose(cval, torch.tensor([one[k] for k in sorted(one)]))


@pytest.mark.parametrize("name", ["ogbl-ddi", "ogbl-collab", "ogbl-citation2"])
def test_synthetic_dataset_has_the_ogb_layout(name, monkeypatch):
    monkeypatch.setenv("PLNLP_SYNTH_SCALE", "0.002")
    ds = shims.SyntheticLinkPropPredDataset(name)
    data, split = ds[0], ds.get_edge_split()
    assert set(split) == {"train", "valid", "test"} and data.edge_index.size(0) == 2
    if name == "ogbl-citation2":
        assert set(split["valid"]) == {"source_node", "target_node", "target_node_neg"}
        assert split["valid"]["target_node_neg"].size(0) == split["valid"]["source_node"].numel()
    else:
        assert split["train"]["edge"].size(1) == 2 and "edge_neg" in split["valid"]
        assert data.edge_index.size(1) == 2 * split["train"]["edge"].size(0)     # both directions, like OGB
    assert (name == "ogbl-collab") == hasattr(data, "edge_weight")
    assert int(data.edge_index.max()) < data.num_nodes


@pytest.mark.skipif(not os.path.exists(MAIN), reason="the reference checkout is only present in the build container")
@pytest.mark.skipif(torch.cuda.is_available(), reason="on a CUDA box main.py would go on to train; covered by the GPU tests")
@pytest.mark.parametrize("argv", [
    ["--data_name", "ogbl-ddi", "--encoder", "SAGE"],
    ["--data_name", "ogbl-collab", "--predictor", "DOT", "--use_valedges_as_input", "True", "--year", "2010",
     "--use_coalesce", "True", "--loss_func", "WeightedHingeAUC"],
    ["--data_name", "ogbl-citation2", "--encoder", "GCN", "--use_node_feats", "True", "--train_node_emb", "True",
     "--eval_metric", "mrr", "--neg_sampler", "local"],
    ["--data_name", "ogbl-collab", "--encoder", "WSAGE"],
    ["--data_name", "ogbl-ddi", "--encoder", "Transformer", "--predictor", "MLPCAT", "--neg_sampler", "global_perm"],
])
def test_reference_main_runs_unchanged_up_to_the_model(argv, tmp_path, monkeypatch):
    """the reference's own main.py, imported through the shims: argument parsing, dataset, ToSparseTensor, the
    per-dataset graph preparation (to_symmetric / to_undirected / coalesce / weight normalisation), data.to(device)
    and gcn / adj normalisation all run; BaseModel then refuses the CUDA-less box"""
    monkeypatch.setenv("PLNLP_SYNTH_SCALE", "0.002")
    saved = dict(sys.modules)
    monkeypatch.setattr(sys, "argv", ["main.py", "--res_dir", str(tmp_path), "--epochs", "1", "--runs", "1"] + argv)
    try:
        shims.install()
        with pytest.raises(RuntimeError, match="CUDA"):
            runpy.run_path(MAIN, run_name="__main__")
    finally:
        for k in [k for k in sys.modules if k not in saved]:
            del sys.modules[k]
        sys.modules.update(saved)
    assert any(f.startswith("log_") for f in os.listdir(tmp_path))      # main.py got as far as writing its log header
