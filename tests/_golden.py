"""Helpers shared by the parity tests: load a golden fixture and rebuild its seeded market."""
import glob
import hashlib
import json
import os

import numpy as np

from incorporating_different_sources_b200.synthetic import generate_market

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_market_cache = {}


def golden_names():
    """Per-window fixtures (the loop-level ``bt_*`` fixtures are handled by test_gpu_backtest.py)."""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith(("bt_", "est_")))


def estimator_golden_names():
    """Fixtures of the sibling estimators (tests/golden/make_estimator_golden.py)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "est_*.npz")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def market_for(meta):
    key = json.dumps(meta["market"], sort_keys=True)
    if key not in _market_cache:
        if len(_market_cache) > 2:
            _market_cache.clear()
        mkt = generate_market(**meta["market"])
        # the fixtures are only meaningful if the seeded generator reproduces the same inputs
        assert sha(mkt.prices) == meta["sha_prices"], "synthetic generator drifted: daily prices differ"
        assert sha(mkt.hf_prices) == meta.get("sha_hf", sha(mkt.hf_prices)), "synthetic generator drifted: intraday prices differ"
        assert sha(mkt.caps) == meta.get("sha_caps", sha(mkt.caps)), "synthetic generator drifted: caps differ"
        _market_cache[key] = mkt
    return _market_cache[key]


def relerr(a, b):
    """max_i |a-b| / max_i |b| (SURVEY §8(d) parity definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def check_matrix(name, M, z, pre, tol):
    """Compare a matrix against a golden stored fully or in compact (diag/rows/fro) form."""
    if pre + name in z:
        assert relerr(M, z[pre + name]) <= tol, name
        return
    ref_diag = z[pre + name + "_diag"]
    ref_rows = z[pre + name + "_rows"]
    scale = np.max(np.abs(ref_diag))
    assert np.max(np.abs(np.diag(M) - ref_diag)) / scale <= tol, name + " diag"
    n = M.shape[0]
    assert np.max(np.abs(M[[0, n // 2, n - 1]] - ref_rows)) / scale <= tol, name + " rows"
    assert abs(np.linalg.norm(M) - float(z[pre + name + "_fro"])) / float(z[pre + name + "_fro"]) <= tol, name + " fro"
