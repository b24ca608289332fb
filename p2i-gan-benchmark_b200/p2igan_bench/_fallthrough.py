"""Fall-through from the alias package to the reference checkout for everything the hot path does not replace.

``p2igan_bench`` here shadows ONLY ``models.{P2IGenerator,P2IDiscriminator,build_*}``, ``modules.{ReconstructionLoss,gan_loss}``
and ``metrics.{MetricConfig,RainfallMetricSuite,...}``.  Every other import the reference's scripts make
(``p2igan_bench.data.dataloader`` at scripts/train.py:20 and scripts/infer.py:16; ``DKGenerator`` / ``STDKGenerator`` at
scripts/infer.py:17; ``p2igan_bench.modules.layer`` from models/dk.py:7; ``p2igan_bench.config``) resolves to the reference's own
files: the package ``__path__`` of every alias (sub)package is extended with the matching directory of the reference checkout,
and names the alias does not define are looked up lazily in the reference's package of the same name.

The reference checkout is found (in this order) through ``$P2I_REFERENCE_ROOT``, any other ``sys.path`` entry that holds a
``p2igan_bench/__init__.py``, and the current working directory (the reference's scripts are run from its root).
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_pkg_dirs() -> List[str]:
    """Directories of OTHER ``p2igan_bench`` packages (the reference checkout), most specific first."""
    cands = []
    env = os.environ.get("P2I_REFERENCE_ROOT")
    if env:
        cands.append(env)
    cands += [p or os.getcwd() for p in sys.path]
    cands.append(os.getcwd())
    out = []
    for root in cands:
        d = os.path.abspath(os.path.join(root, "p2igan_bench"))
        if d != _HERE and d not in out and os.path.isfile(os.path.join(d, "__init__.py")):
            out.append(d)
    return out


def extended_path(own_path, sub: str = "") -> List[str]:
    """``__path__`` for the alias (sub)package: our directory first, then the reference's directory of the same name."""
    path = list(own_path)
    for d in reference_pkg_dirs():
        r = os.path.join(d, sub) if sub else d
        if os.path.isdir(r) and r not in path:
            path.append(r)
    return path


def reference_module(sub: str):
    """The reference's ``p2igan_bench.<sub>`` package, loaded under the private name ``p2igan_bench.<sub>._reference`` so that
    its relative imports (``from .dk import DKGenerator``) resolve to the reference's files.  None when no checkout is found."""
    name = f"p2igan_bench.{sub}._reference"
    if name in sys.modules:
        return sys.modules[name]
    for d in reference_pkg_dirs():
        init = os.path.join(d, sub, "__init__.py")
        if os.path.isfile(init):
            spec = importlib.util.spec_from_file_location(name, init, submodule_search_locations=[os.path.join(d, sub)])
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                sys.modules.pop(name, None)
                raise
            return mod
    return None


def reference_attr(sub: str, attr: str, alias_name: str):
    """Module-level ``__getattr__`` body for an alias (sub)package."""
    mod: Optional[object] = reference_module(sub)
    if mod is not None and hasattr(mod, attr):
        return getattr(mod, attr)
    where = "the reference checkout was not found (set P2I_REFERENCE_ROOT or put it on sys.path after this package)" \
        if mod is None else "the reference does not define it either"
    raise AttributeError(f"module '{alias_name}' has no attribute '{attr}': it is outside the sm_100a hot path and {where}")
