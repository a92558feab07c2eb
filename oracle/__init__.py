"""CPU oracle for the gnn-mlp hot path — TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under the product package does (it has no CPU path at all).

What it restates (plain PyTorch fp32 / fp64 and numpy, straightforward per-edge formulation):
  * ``ptgnn`` message passing + ``torch_scatter`` segment ops — the arithmetic of the path.  Those two
    packages are UNPINNED third-party dependencies of the reference (requirements.txt:13,
    Dockerfile:9,14) that are absent from /root/reference and not installable offline, and the
    reference has no test, golden vector or fixture at that boundary, so for these
    pieces the oracle is **parity unpinned**: it follows the published semantics of
    ptgnn.MlpMessagePassingLayer / torch_scatter (CPU) restated in SURVEY.md §8a P1-P7 and is pinned only
    by the hand-computed known-answer tests in tests/test_oracle_kat.py.
  * the in-repo pieces (heads, detector and selector/generator losses, rewrite bookkeeping, data schema) — restated with file:line
    citations and pinned against the REAL reference code imported from /root/reference
    (tests/golden/make_golden.py wrote the fixtures under tests/golden/).
  * ``seq_ref.py`` — the relational-transformer layer of the seq-great / seq-rat models (torch-only
    reference files, so fully pinned: outputs and gradients of the real classes in tests/golden/seq_layers.npz);
    ``seq_model_ref.py`` — the whole sequence module on top of it, pinned by tests/golden/seq_model.npz.
"""
