"""Drop-in alias package: the import paths the reference's scripts use (``p2igan_bench.models``, ``.modules``,
``.metrics``) resolved to the sm_100a implementation in ``p2igan_b200``.  Put ``p2i-gan-benchmark_b200/`` ahead of the
reference checkout on ``sys.path`` and ``scripts/train.py`` / ``scripts/infer.py`` pick up these modules for the hot path
(see INTEGRATION.md)."""
