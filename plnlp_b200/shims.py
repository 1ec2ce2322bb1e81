"""Import shims: let the reference's ``main.py`` run UNCHANGED on a machine that has neither PyG, torch_sparse,
torch_cluster nor ogb.

``install()`` registers stand-ins in ``sys.modules`` for exactly the names ``/root/reference/main.py:6-13``
imports:

    torch_geometric.transforms.ToSparseTensor     -> ``ToSparseTensor``      (adjacency = graph.CSRGraph)
    torch_geometric.utils.to_undirected           -> ``to_undirected``
    torch_sparse.SparseTensor / coalesce          -> ``SparseTensor`` / ``coalesce``
    torch_cluster.random_walk                     -> augment.random_walk      (GPU kernel)
    ogb.linkproppred.PygLinkPropPredDataset       -> ``SyntheticLinkPropPredDataset`` (OGB SHAPES, synthetic data:
                                                     there is no network / dataset on the boxes this runs on)
    ogb.linkproppred.Evaluator                    -> ``Evaluator``            (ranking on the GPU)
    plnlp, plnlp.model, plnlp.utils, ...          -> plnlp_b200.*

Everything here is index plumbing in torch (any device); the arithmetic stays in the kernels.  A module that is
really installed is never replaced.  Status: the data-preparation half of ``main.py`` is exercised on the CPU
against the oracle's restatement of the same third-party calls (tests/test_shims_cpu.py, which also executes the
real ``main.py`` up to the point where it constructs ``BaseModel`` on a CUDA-less box); the training half is the
``BaseModel`` surface the GPU tests cover.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

from .graph import CSRGraph


# ---------------------------------------------------------------------------------------------- torch_sparse
def SparseTensor(row=None, rowptr=None, col=None, value=None, sparse_sizes=None, is_sorted=False):
    """``torch_sparse.SparseTensor(row=, col=, value=)`` as main.py:124-126,137-139,184 calls it."""
    if row is not None:
        return CSRGraph.from_coo(row, col, value, sparse_sizes, is_sorted=is_sorted)
    return CSRGraph(rowptr, col, value, sparse_sizes)


def coalesce(index, value, m, n, op="add"):
    """``torch_sparse.coalesce`` (main.py:142): sort by (row, col), merge duplicates (values summed)."""
    if op not in ("add", "sum"):
        raise NotImplementedError(f"coalesce op={op!r}")
    key = index[0].to(torch.int64) * int(n) + index[1].to(torch.int64)
    ukey, inv = torch.unique(key, return_inverse=True)
    out_index = torch.stack([torch.div(ukey, int(n), rounding_mode="floor"), ukey % int(n)])
    if value is None:
        return out_index, None
    out_value = torch.zeros((ukey.numel(),) + tuple(value.shape[1:]), dtype=value.dtype, device=value.device)
    out_value.index_add_(0, inv, value)
    return out_index, out_value


# ---------------------------------------------------------------------------------------------- torch_geometric
def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    """``torch_geometric.utils.to_undirected`` (2.0.1 call shape, main.py:122,135): both directions, coalesced;
    returns ``(edge_index, edge_attr)`` when an attribute is given, else ``edge_index``."""
    if isinstance(edge_attr, int):                    # (edge_index, num_nodes) legacy call
        edge_attr, num_nodes = None, edge_attr
    n = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    attr = None if edge_attr is None else torch.cat([edge_attr, edge_attr], dim=0)
    idx, val = coalesce(both, attr, n, n, op=reduce)
    return idx if edge_attr is None else (idx, val)


class Data:
    """the slice of ``torch_geometric.data.Data`` main.py touches: attribute bag, ``to(device)``,
    ``num_features`` (0 without ``x``), ``num_nodes``"""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_features(self):
        x = getattr(self, "x", None)
        return 0 if x is None else (1 if x.dim() == 1 else x.size(-1))

    def to(self, device, *args, **kwargs):
        out = Data()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.to(device) if hasattr(v, "to") else v
        return out

    def __contains__(self, key):
        return key in self.__dict__


class ToSparseTensor:
    """``T.ToSparseTensor()`` (main.py:81): ``adj_t[dst, src]`` sorted by (dst, src), value = ``edge_weight`` when
    present, duplicates kept; ``edge_index`` / ``edge_weight`` are removed (the transform's default)."""

    def __init__(self, attr="edge_weight", remove_edge_index=True, fill_cache=True):
        self.attr, self.remove_edge_index = attr, remove_edge_index

    def __call__(self, data):
        w = getattr(data, self.attr, None) if self.attr else None
        data.adj_t = CSRGraph.from_edge_index(data.edge_index, w, data.num_nodes)
        if self.remove_edge_index:
            del data.edge_index
            if w is not None:
                delattr(data, self.attr)
        return data


# ---------------------------------------------------------------------------------------------- ogb
class Evaluator:
    """``ogb.linkproppred.Evaluator``: ``BaseModel.test`` ranks on the GPU and only sets ``K`` on this object
    (utils.py:49-56); ``eval`` is kept for code that calls it directly."""

    def __init__(self, name):
        self.name = name
        self.eval_metric = "mrr" if "citation" in name else "hits"
        self.K = {"ogbl-ddi": 20, "ogbl-collab": 50, "ogbl-ppa": 100}.get(name, 20)

    def eval(self, input_dict):
        from . import utils
        pos, neg = input_dict["y_pred_pos"], input_dict["y_pred_neg"]
        if self.eval_metric == "mrr":
            return {"mrr_list": utils.mrr_list(utils._cuda_f32(pos), utils._cuda_f32(neg).reshape(pos.numel(), -1))}
        return {f"hits@{self.K}": utils.hits_at_k(utils._cuda_f32(pos), utils._cuda_f32(neg), self.K)}


# OGB shapes: nodes, train edges, features, eval sizes
_SHAPES = {
    "ogbl-ddi": dict(N=4267, E=1067911, feats=0, directed=False, n_eval=133489, n_eval_neg=101882),
    "ogbl-collab": dict(N=235868, E=1179052, feats=128, directed=False, n_eval=60084, n_eval_neg=100000,
                        weighted=True),
    "ogbl-citation2": dict(N=2927963, E=30387995, feats=128, directed=True, n_eval=86596, n_eval_neg=1000),
}


class SyntheticLinkPropPredDataset:
    """Stand-in for ``PygLinkPropPredDataset(name=, root=)`` with the SHAPES of the OGB link-property datasets
    and synthetic content (seeded; ``PLNLP_SYNTH_SCALE`` < 1 shrinks node and edge counts for quick runs).
    ``dataset[0]`` -> Data(edge_index [2, E'], x, edge_weight / edge_year for collab, num_nodes);
    ``get_edge_split()`` -> the split dictionaries ``plnlp/utils.py:7-41`` reads."""

    def __init__(self, name, root=None, seed=0):
        if name not in _SHAPES:
            raise ValueError(f"no synthetic shape for {name!r}")
        self.name = name
        cfg = dict(_SHAPES[name])
        scale = float(os.environ.get("PLNLP_SYNTH_SCALE", "1"))
        N = max(int(cfg["N"] * scale), 64)
        E = max(int(cfg["E"] * scale), 4 * N if scale < 1 else 1)
        E = min(E, N * (N - 1) // 4)
        g = torch.Generator().manual_seed(seed)
        src = torch.randint(0, N, (int(E * 1.3),), generator=g)
        dst = torch.randint(0, N, (int(E * 1.3),), generator=g)
        keep = src != dst
        src, dst = src[keep], dst[keep]
        if not cfg["directed"]:
            key = torch.unique(torch.minimum(src, dst) * N + torch.maximum(src, dst))
            key = key[torch.randperm(key.numel(), generator=g)[:E]]
            src, dst = torch.div(key, N, rounding_mode="floor"), key % N
        else:
            src, dst = src[:E], dst[:E]
        self._train = torch.stack([src, dst], 1)
        n_ev = max(int(cfg["n_eval"] * scale), 16)
        n_neg = cfg["n_eval_neg"] if cfg["directed"] else max(int(cfg["n_eval_neg"] * scale), 16)
        if cfg["directed"]:
            n_neg = max(int(n_neg * min(1.0, scale * 50)), 8)
        x = torch.randn(N, cfg["feats"], generator=g) if cfg["feats"] else None
        data = Data(num_nodes=N, x=x)
        if cfg["directed"]:
            data.edge_index = self._train.t().contiguous()
        else:
            data.edge_index = torch.cat([self._train.t(), self._train.t().flip(0)], 1).contiguous()
        self._split = {}
        if cfg.get("weighted"):
            w = torch.randint(1, 6, (self._train.size(0),), generator=g)
            year = torch.randint(1990, 2018, (self._train.size(0),), generator=g)
            data.edge_weight = torch.cat([w, w]).reshape(-1, 1)
            data.edge_year = torch.cat([year, year]).reshape(-1, 1)
            self._split["train"] = {"edge": self._train, "weight": w, "year": year}
        elif cfg["directed"]:
            self._split["train"] = {"source_node": src.contiguous(), "target_node": dst.contiguous()}
        else:
            self._split["train"] = {"edge": self._train}
        for part in ("valid", "test"):
            if cfg["directed"]:
                self._split[part] = {"source_node": torch.randint(0, N, (n_ev,), generator=g),
                                     "target_node": torch.randint(0, N, (n_ev,), generator=g),
                                     "target_node_neg": torch.randint(0, N, (n_ev, n_neg), generator=g)}
            else:
                rec = {"edge": torch.randint(0, N, (n_ev, 2), generator=g),
                       "edge_neg": torch.randint(0, N, (n_neg, 2), generator=g)}
                if cfg.get("weighted"):
                    rec["weight"] = torch.randint(1, 6, (n_ev,), generator=g)
                    rec["year"] = torch.full((n_ev,), 2018 if part == "valid" else 2019)
                self._split[part] = rec
        self._data = data

    def __getitem__(self, i):
        if i != 0:
            raise IndexError(i)
        return self._data

    def __len__(self):
        return 1

    def get_edge_split(self):
        return {k: dict(v) for k, v in self._split.items()}


# ---------------------------------------------------------------------------------------------- install
def _random_walk(row, col, start, walk_length, *args, **kwargs):
    from . import augment
    return augment.random_walk(row, col, start, walk_length, *args, **kwargs)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__plnlp_b200_shim__ = True
    return m


def _really_installed(name):
    try:
        mod = importlib.import_module(name)
    except Exception:
        return False
    return not getattr(mod, "__plnlp_b200_shim__", False)


def install(alias_plnlp=True):
    """register the stand-ins (see the module docstring); returns the list of module names shimmed"""
    from . import layer
    done = []

    def put(name, mod):
        sys.modules[name] = mod
        done.append(name)

    if not _really_installed("torch_sparse"):
        put("torch_sparse", _module("torch_sparse", SparseTensor=SparseTensor, coalesce=coalesce))
    if not _really_installed("torch_geometric"):
        tr = _module("torch_geometric.transforms", ToSparseTensor=ToSparseTensor)
        ut = _module("torch_geometric.utils", to_undirected=to_undirected)
        nn = _module("torch_geometric.nn", SAGEConv=layer.SAGEConv, GCNConv=layer.GCNConv, GraphConv=layer.GraphConv,
                     TransformerConv=layer.TransformerConv)
        da = _module("torch_geometric.data", Data=Data)
        put("torch_geometric", _module("torch_geometric", transforms=tr, utils=ut, nn=nn, data=da))
        for sub, mod in (("transforms", tr), ("utils", ut), ("nn", nn), ("data", da)):
            put("torch_geometric." + sub, mod)
    if not _really_installed("torch_cluster"):
        put("torch_cluster", _module("torch_cluster", random_walk=_random_walk))
    if not _really_installed("ogb"):
        lp = _module("ogb.linkproppred", PygLinkPropPredDataset=SyntheticLinkPropPredDataset, Evaluator=Evaluator)
        put("ogb", _module("ogb", linkproppred=lp))
        put("ogb.linkproppred", lp)
    if alias_plnlp:
        import plnlp_b200
        put("plnlp", plnlp_b200)
        for sub in ("logger", "model", "utils", "layer", "loss", "negative_sample"):
            put("plnlp." + sub, importlib.import_module("plnlp_b200." + sub))
    return done
