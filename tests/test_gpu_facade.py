"""The drop-in facade (reference function names / signatures, pandas in -> pandas out) on the GPU,
against the reference-generated golden vectors.  The calls below are written the way the reference's
own dispatcher makes them (portfolio_calculations.py:954-1034), including the weekly-window cases."""
import numpy as np
import pandas as pd
import pytest

from tests._golden import check_matrix, golden_names, load_golden, market_for, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9
D_TO_FREQ = {1: "daily", 7: "weekly", 31: "monthly"}
SMALL = [n for n in golden_names() if "n500" not in n]


def _frames(mkt, md, names, d):
    caps_df = md["stock_market_caps_df"][names].loc[:d]
    prices_df = md["stock_prices_df"][names].loc[:d]
    d_incl = d.replace(hour=23, minute=59, second=59)
    intr_df = md["stock_intraday_prices_df"][names]
    intr_df = intr_df.loc[intr_df.index <= d_incl]
    return caps_df, prices_df, intr_df, md["risk_free_rate_df"]


@pytest.mark.parametrize("name", SMALL)
def test_facade_matches_reference(name):
    from incorporating_different_sources_b200 import portfolio_calculations as pc
    z, meta = load_golden(name)
    mkt = market_for(meta)
    md = mkt.market_data()
    spec = meta["spec"]
    conj = spec["weighting_strategy"].startswith("conjugate")
    for wi, w in enumerate(meta["windows"]):
        pre = f"w{wi}_"
        d = pd.Timestamp(mkt.dates[w["d_idx"]])
        names = [mkt.tickers[c] for c in z[pre + "cols"]]
        caps_df, prices_df, intr_df, rf_df = _frames(mkt, md, names, d)
        t_df = pc.calculate_canonical_statistics_t(spec, d, prices_df, rf_df)
        T_df = pc.calculate_canonical_statistics_T(spec, d, prices_df, rf_df)
        assert list(t_df.index) == names and list(T_df.index) == names and list(T_df.columns) == names
        assert list(t_df.columns) == [0]
        assert relerr(t_df.values[:, 0], z[pre + "t"]) <= TOL
        check_matrix("T", T_df.values, z, pre, TOL)
        if not conj:
            nu_df = pc.calculate_mean_jeffreys_posterior_nu(spec, d, prices_df, rf_df)
            w_df = pc.calculate_jeffreys_portfolio(spec, d, prices_df, rf_df)
            assert list(w_df.index) == names and list(w_df.columns) == ["Weight"]
            assert relerr(nu_df["Weight"].values, z[pre + "nu"]) <= TOL
            assert relerr(w_df["Weight"].values, z[pre + "weights"]) <= TOL
            continue
        mcm_key = "vix_prices_df" if "vix" in spec["weighting_strategy"] else "epu_prices_df"
        mcm_df = md[mcm_key].loc[md[mcm_key].index <= d]
        n0 = pc.calculate_conjugate_prior_n(spec, d, mcm_df)
        n1 = pc.calculate_conjugate_posterior_n(spec, d, mcm_df)
        assert abs(n0 - float(z[pre + "n0"])) <= TOL * n0 and abs(n1 - float(z[pre + "n1"])) <= TOL * n1
        hf_days = meta["hf_days"]
        if hf_days is None:
            S0_df = pc.calculate_conjugate_prior_S(spec, d, intr_df, mcm_df)
        else:
            S0_df = pc.calculate_conjugate_prior_S(dict(spec, rolling_window_frequency=D_TO_FREQ[hf_days]), d,
                                                   intr_df, mcm_df, conjugate_prior_n=n0)
        check_matrix("S0", S0_df.loc[names, names].values, z, pre, TOL)
        w0_df = pc.calculate_conjugate_prior_w(spec, d, prices_df, caps_df, mcm_df)
        assert w0_df.index.name == "Stock" and relerr(w0_df["Weight"].reindex(names).values, z[pre + "w0"]) <= TOL
        c = pc.calculate_conjugate_c(spec, d, prices_df, caps_df, intr_df, mcm_df, conjugate_prior_S_df=S0_df)
        assert abs(c - float(z[pre + "c"])) <= TOL * abs(c)
        S1_df = pc.calculate_conjugate_posterior_S(spec, d, prices_df, intr_df, mcm_df, rf_df,
                                                   conjugate_prior_S_df=S0_df)
        check_matrix("S1", S1_df.loc[names, names].values, z, pre, TOL)
        w1_df = pc.calculate_conjugate_posterior_w(spec, d, prices_df, caps_df, intr_df, mcm_df, rf_df,
                                                   conjugate_c=c, conjugate_prior_S_df=S0_df,
                                                   conjugate_posterior_S_df=S1_df)
        assert list(w1_df.index) == names
        assert relerr(w1_df["Weight"].values, z[pre + "w1"]) <= TOL
        v1 = pc.calculate_portfolio_variance(w1_df, S1_df)
        assert abs(v1 - float(z[pre + "v1"])) <= TOL * abs(v1)
        nu_df = pc.calculate_mean_conjugate_posterior_nu(spec, d, prices_df, caps_df, intr_df, mcm_df, rf_df,
                                                         conjugate_c=c, conjugate_prior_S_df=S0_df,
                                                         conjugate_posterior_S_df=S1_df)
        assert relerr(nu_df["Weight"].values, z[pre + "nu"]) <= TOL
        nu2_df = pc.calculate_mean_conjugate_posterior_nu(spec, d, prices_df, caps_df, intr_df, mcm_df, rf_df,
                                                          conjugate_prior_S_df=S0_df, conjugate_posterior_w_df=w1_df)
        assert relerr(nu2_df["Weight"].values, z[pre + "nu"]) <= TOL
        if hf_days is None:
            w_df = pc.calculate_conjugate_hf_mcm_portfolio(spec, d, caps_df, prices_df, intr_df, mcm_df, rf_df)
            assert list(w_df.index) == names and list(w_df.columns) == ["Weight"] and w_df.index.name is None
            assert relerr(w_df["Weight"].values, z[pre + "weights"]) <= TOL
            S1f = pc.calculate_conjugate_posterior_S(spec, d, prices_df, intr_df, mcm_df, rf_df)
            check_matrix("S1", S1f.loc[names, names].values, z, pre, TOL)
            w1f = pc.calculate_conjugate_posterior_w(spec, d, prices_df, caps_df, intr_df, mcm_df, rf_df)
            assert relerr(w1f["Weight"].values, z[pre + "w1"]) <= TOL


def test_facade_excess_returns_and_errors():
    from incorporating_different_sources_b200 import portfolio_calculations as pc
    from oracle import bayes_oracle as bo
    z, meta = load_golden("c1_n10_conj_vix_vw")
    mkt = market_for(meta)
    md = mkt.market_data()
    spec = meta["spec"]
    d = pd.Timestamp(mkt.dates[-1])
    win = md["stock_prices_df"].iloc[-252:]
    X = pc.calculate_excess_log_returns_from_prices(spec, win, md["risk_free_rate_df"])
    ref = bo.excess_log_returns(mkt.dates[-252:], mkt.prices[-252:], mkt.dates, mkt.rf)
    assert X.shape == (251, 10) and list(X.columns) == mkt.tickers and X.index[0] == win.index[1]
    assert relerr(X.values, ref) <= 1e-12
    # error conventions of the reference
    with pytest.raises(ValueError):        # last date mismatch, :145-147
        pc.calculate_canonical_statistics_T(spec, d, md["stock_prices_df"].iloc[:-1], md["risk_free_rate_df"])
    with pytest.raises(ValueError):        # MCM last date mismatch, :98-100
        pc.calculate_conjugate_prior_n(spec, d, md["vix_prices_df"].iloc[:-2])
    with pytest.raises(RuntimeError):      # unknown window frequency, :308
        pc.calculate_conjugate_prior_S(dict(spec, rolling_window_frequency="hourly"), d,
                                       md["stock_intraday_prices_df"], md["vix_prices_df"], conjugate_prior_n=1.0)
    with pytest.raises(ValueError):        # unknown prior weights, :378
        pc.calculate_conjugate_prior_w(dict(spec, weighting_strategy="conjugate_hf_vix_xx"), d,
                                       md["stock_prices_df"], md["stock_market_caps_df"], md["vix_prices_df"])
    gap = md["stock_prices_df"].iloc[-252:].copy()
    gap.index = gap.index[:-1].append(pd.DatetimeIndex([gap.index[-1] + pd.Timedelta(days=30)]))
    with pytest.raises(AssertionError):    # date gap, :44
        pc.calculate_canonical_statistics_t(spec, gap.index[-1], gap, md["risk_free_rate_df"])
