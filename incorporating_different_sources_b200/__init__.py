"""B200-native rolling-window Bayesian tangency-portfolio weights.

A drop-in accelerator for one hot path of ``vilnik/incorporating-different-sources``
(``src/portfolio_calculations.py:31-608,819-849`` as driven by the backtest loop): hand-written
sm_100a CUDA kernels behind a C-ABI (``include/bayes_portfolio.h``), a batched host engine and a
facade with the reference's own function signatures (``portfolio_calculations`` in this package).
"""
from .synthetic import SyntheticMarket, generate_market  # noqa: F401
from .windows import WindowBatch, plan_daily_windows  # noqa: F401

__all__ = ["SyntheticMarket", "generate_market", "WindowBatch", "plan_daily_windows", "BayesEngine"]


def __getattr__(name):
    if name == "BayesEngine":
        from .engine import BayesEngine
        return BayesEngine
    raise AttributeError(name)
