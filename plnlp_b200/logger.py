"""Run bookkeeping with the interface of /root/reference/plnlp/logger.py (pure host code, not on
the hot path): per-run lists of (valid, test) results, best-valid selection, mean/std over runs."""
from __future__ import annotations

import sys

import torch


class Logger(object):
    def __init__(self, runs, info=None):
        self.info = info
        self.results = [[] for _ in range(runs)]

    def add_result(self, run, result):
        assert len(result) == 2
        assert 0 <= run < len(self.results)
        self.results[run].append(result)

    @staticmethod
    def _best(r, last_best):
        valid = r[:, 0]
        if last_best:  # index of the LAST maximum
            idx = valid.numel() - 1 - int(valid.flip(0).argmax())
        else:
            idx = int(valid.argmax())
        return float(valid.max()), idx, float(r[idx, 1])

    def print_statistics(self, run=None, f=sys.stdout, last_best=False):
        if run is not None:
            best_valid, idx, test = self._best(100 * torch.tensor(self.results[run]), last_best)
            print(f'Run {run + 1:02d}:', file=f)
            print(f'Highest Valid: {best_valid:.2f}', file=f)
            print(f'Highest Eval Point: {idx + 1}', file=f)
            print(f'   Final Test: {test:.2f}', file=f)
            return
        picks = [self._best(100 * torch.tensor(r), last_best) for r in self.results]
        best = torch.tensor([[v, t] for v, _, t in picks])
        print('All runs:', file=f)
        print(f'Highest Valid: {best[:, 0].mean():.2f}  {best[:, 0].std():.2f}', file=f)
        print(f'   Final Test: {best[:, 1].mean():.2f}  {best[:, 1].std():.2f}', file=f)
