"""Shared helpers for the parity tests."""
import torch


def rel_err(a, b):
    """max|a-b| / max(|b|, eps): the per-tensor relative error the 1e-5 bar of
    BASELINE.json's north_star is checked with."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def rand_graph(N, E, seed, weighted=False, hub=False):
    """Random directed multigraph-free edge list [2, E'] incl. isolated nodes,
    optional hub row; returns (edge_index, weight)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g)
    dst = torch.randint(0, max(N - 3, 1), (E,), generator=g)  # last rows stay empty
    if hub and N > 4:
        hub_src = torch.arange(N)
        src = torch.cat([src, hub_src, hub_src])
        dst = torch.cat([dst, torch.full((N,), 2), torch.full((N,), 2)])  # duplicates kept
    ei = torch.stack([src, dst])
    w = torch.rand(ei.size(1), generator=g) + 0.5 if weighted else None
    return ei, w


def map_encoder_state(state):
    """reference module state_dict (convs.i.*) -> oracle names (enc.i.*)"""
    return {"enc." + k[len("convs."):]: v for k, v in state.items()}


def map_predictor_state(state):
    return {"pred." + k[len("lins."):]: v for k, v in state.items()}


def fp32_close(got, ref32, ref64, tol=1e-5, slack=8.0, floor=0.0):
    """Parity criterion for ill-conditioned quantities (sums with heavy cancellation, e.g. the bias
    gradient of a pairwise loss whose d loss/d score sums to exactly zero): pass when the result is
    within ``tol`` relative of the exact (fp64) value, OR no further from it than ``slack`` times the
    reference's own fp32 rounding error.  Returns (ok, message)."""
    got, ref32, ref64 = (t.detach().double().cpu().reshape(-1) for t in (got, ref32, ref64))
    scale = max(float(ref64.abs().max()), 1e-30)
    e_got = float((got - ref64).abs().max())
    e_ref = float((ref32 - ref64).abs().max())
    # ``floor``: absolute error allowed regardless (tol x the magnitude of the LARGEST gradient tensor
    # of the model) -- for quantities whose exact value is 0 by symmetry (sum of d loss/d score)
    ok = e_got <= tol * scale or e_got <= slack * e_ref + 1e-12 * scale or e_got <= floor
    return ok, f"err {e_got:.3e} (rel {e_got / scale:.2e}), reference fp32 err {e_ref:.3e}, scale {scale:.3e}"
