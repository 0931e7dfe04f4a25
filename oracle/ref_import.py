"""TEST INFRASTRUCTURE — loader for the UNMODIFIED reference hot path.

Imports ``/root/reference/src/portfolio_calculations.py`` as it lies (nothing is
copied into this repo) so that (i) the NumPy restatement in
``oracle/bayes_oracle.py`` can be pinned against the reference itself and (ii)
``tests/golden/make_golden.py`` can generate the committed golden vectors.

The reference cannot be imported plainly:
* ``portfolio_calculations.py:4-8`` imports ``pypfopt`` (not installed; used only
  by the out-of-scope shrinkage / Black-Litterman strategies);
* ``portfolio_calculations.py:12`` imports ``data_handling``, whose vendor
  sub-modules need API keys and ``os.makedirs`` into the read-only tree.
Two stub modules are registered in ``sys.modules`` before the import; the only
stub function ever called on the hot path is ``extract_unique_tickers``
(``portfolio_calculations.py:619``).

Sources are looked up at ``$REF_SRC``, ``/root/reference/src`` (build container) and ``baseline/_ref/src`` (the
git-ignored staging copy ``__graft_entry__.build()`` makes, which travels to the GPU box).  The ``-m gpu`` tests
and ``smoke()`` never call it; ``bench.py`` does so only for its CPU arm (``--impl reference`` / ``cpu_baseline``).
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# /root/reference exists in the build container only; __graft_entry__.build() stages the unmodified sources under
# baseline/_ref/src (git-ignored, travels to the GPU box with the snapshot) for bench.py's CPU arm
REF_SRC_CANDIDATES = [os.environ.get("REF_SRC", ""), "/root/reference/src", os.path.join(_REPO, "baseline", "_ref", "src")]

_state = {"module": None, "tickers": []}


def reference_available() -> bool:
    return any(p and os.path.isfile(os.path.join(p, "portfolio_calculations.py")) for p in REF_SRC_CANDIDATES)


def _placeholder(name):
    def _raise(*a, **k):  # pragma: no cover - never on the hot path
        raise RuntimeError(f"pypfopt.{name} is a stub: strategy out of scope")
    return _raise


def set_universe(tickers):
    """Tickers returned by the stubbed ``data_handling.extract_unique_tickers``."""
    _state["tickers"] = list(tickers)


def load_reference(check: bool = False):
    """Import (once) and return the reference ``portfolio_calculations`` module."""
    if _state["module"] is not None:
        _state["module"].CHECK = check
        return _state["module"]
    src = next((p for p in REF_SRC_CANDIDATES
                if p and os.path.isfile(os.path.join(p, "portfolio_calculations.py"))), None)
    if src is None:
        raise FileNotFoundError("reference sources not found (expected /root/reference/src)")
    os.environ.setdefault("LOGGING_LEVEL", "WARNING")

    pyp = types.ModuleType("pypfopt")
    pyp.EfficientFrontier = _placeholder("EfficientFrontier")
    for sub, names in {
        "risk_models": ["CovarianceShrinkage"],
        "expected_returns": ["mean_historical_return"],
        "black_litterman": ["BlackLittermanModel", "market_implied_prior_returns"],
    }.items():
        m = types.ModuleType(f"pypfopt.{sub}")
        for n in names:
            setattr(m, n, _placeholder(n))
        setattr(pyp, sub, m)
        sys.modules[f"pypfopt.{sub}"] = m
    sys.modules["pypfopt"] = pyp

    dh = types.ModuleType("data_handling")
    dh.extract_unique_tickers = lambda start, end: list(_state["tickers"])
    sys.modules["data_handling"] = dh

    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "reference_portfolio_calculations", os.path.join(src, "portfolio_calculations.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.CHECK = check
    import logging
    mod.logger.setLevel(logging.WARNING)
    _state["module"] = mod
    return mod
