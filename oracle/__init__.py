"""CPU oracle for the PLNLP training / scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the timed CPU baseline), never as the path being measured.
The product package ``plnlp_b200`` never imports this package and has no CPU
fallback.

What it restates
----------------
The reference (``/root/reference``, 8 Python files) delegates all arithmetic to
third-party libraries that are NOT vendored and NOT installable here
(torch_geometric 2.0.1, torch_sparse, torch_cluster, ogb 1.3.2 -- pinned only
by ``README.md:15-19``).  The oracle therefore has two layers:

* ``oracle.sparse`` / ``oracle.pyg`` / ``oracle.ogb_eval`` / ``oracle.sampling``
  restate the *published behaviour* of those third-party pieces (CSR build,
  ``matmul(reduce=...)``, ``SAGEConv`` / ``GCNConv``, ``negative_sampling``,
  ``Evaluator``).  PARITY OF THIS LAYER IS UNPINNED by the reference (it has no
  tests or golden vectors); it is cross-checked against scipy.sparse, fp64
  autograd gradcheck and brute-force loops in ``tests/test_oracle.py``.
* ``oracle.plnlp_ref`` restates the reference's own files (``plnlp/layer.py``,
  ``loss.py``, ``negative_sample.py``, ``utils.py``, ``model.py``).  THIS layer is
  pinned: ``tests/golden/make_golden.py`` imports the real reference package from
  ``/root/reference`` (with the third-party imports satisfied by the layer
  above), runs it on seeded inputs and commits the outputs as fixtures.

``oracle/spmm_ref.c`` is a plain-C restatement of the torch_sparse CPU SpMM loop
(row-parallel, strictly in-order fp32 accumulation per output element); it is
the bit-level reference for summation order and the CPU baseline for the SpMM
micro-benchmark.  Build it with ``make -C oracle``.
"""
