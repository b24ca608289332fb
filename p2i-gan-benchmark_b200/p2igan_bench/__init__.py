"""Drop-in alias package: the import paths the reference's scripts use (``p2igan_bench.models``, ``.modules``,
``.metrics``) resolved to the sm_100a implementation in ``p2igan_b200``.  Put ``p2i-gan-benchmark_b200/`` ahead of the
reference checkout on ``sys.path`` and ``scripts/train.py`` / ``scripts/infer.py`` run UNMODIFIED: the hot path comes from
here, everything else (``p2igan_bench.data``, ``.config``, the DeepKriging / toy models, ``modules.layer`` ...) falls
through to the reference's own files (see ``_fallthrough.py`` and INTEGRATION.md)."""
from ._fallthrough import extended_path as _extended_path

__path__ = _extended_path(__path__)
