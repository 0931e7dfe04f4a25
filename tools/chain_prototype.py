"""CPU prototype of the Jeffreys chain (jeffreys_chain.cu): accuracy of the Woodbury solve relative to a factorised base
window against the reference's arithmetic (np.linalg.inv(J) @ t) and against a per-window Cholesky solve, over a range
of condition numbers.  Test infrastructure (uses the oracle); the numbers are quoted in DESIGN.md 4.8.

    python tools/chain_prototype.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from incorporating_different_sources_b200.synthetic import generate_market
from oracle import bayes_oracle as bo
import scipy.linalg as sl
def run(N,n,G=8,seed=2003):
    mkt=generate_market(N, n+40, seed=seed, bars_per_day=2)
    spec=dict(weighting_strategy="jeffreys",size=N,risk_aversion=5,rolling_window=n,rolling_window_frequency="daily")
    cols=np.arange(N)
    d0=n+5
    L=np.log(mkt.prices[1:]/mkt.prices[:-1])
    def parts(d):
        t,T,X=bo.daily_statistics(spec,mkt,d,cols)
        Lw=L[d-n+1:d]
        a=Lw[:,0]-X[:,0]
        p=Lw.T@a-0.5*(a@a)
        return t,T,p
    t0,T0,p0=parts(d0)
    J0=T0-np.outer(t0,t0)/n
    c=sl.cho_factor(J0,lower=True)
    worst=0; worst_plain=0
    for k in range(1,G):
        d=d0+k
        tk,Tk,pk=parts(d)
        Jk=Tk-np.outer(tk,tk)/n
        ref=np.dot(np.linalg.inv(Jk),tk)          # the reference's arithmetic
        plain=sl.cho_solve(sl.cho_factor(Jk,lower=True),tk)
        news=[L[d0+i-1] for i in range(1,k+1)]
        olds=[L[d0-n+1+i-1] for i in range(1,k+1)]
        U=np.array(news+olds+[pk-p0,np.ones(N),tk,t0]).T
        m=U.shape[1]
        Cinv=np.zeros((m,m))
        for i in range(k): Cinv[i,i]=1.0; Cinv[k+i,k+i]=-1.0
        Cinv[2*k,2*k+1]=Cinv[2*k+1,2*k]=-1.0
        Cinv[2*k+2,2*k+2]=-float(n); Cinv[2*k+3,2*k+3]=float(n)
        Z=sl.solve_triangular(c[0],U,lower=True)
        Wm=Z.T@Z
        Y=sl.solve_triangular(c[0],Z,lower=True,trans='T')
        yt=Y[:,2*k+2]
        M=Cinv+Wm
        x=yt-Y@np.linalg.solve(M,Wm[:,2*k+2])
        e=np.max(np.abs(x-ref))/np.max(np.abs(ref)); worst=max(worst,e)
        ep=np.max(np.abs(plain-ref))/np.max(np.abs(ref)); worst_plain=max(worst_plain,ep)
    return np.linalg.cond(J0), worst, worst_plain
for N,n in ((100,252),(200,252),(230,252),(240,252),(245,252),(248,252),(400,420),(500,520),(500,560),(500,1008)):
    c,w,wp=run(N,n)
    print(f"N={N} n={n} cond(J)={c:.3g} chain err={w:.2e} plain-chol err={wp:.2e}")
