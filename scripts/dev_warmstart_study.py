"""Developer study (CPU, numpy float64): rounds of the batched FISTA driver (same trial / accept logic as
csrc/fista.cu) from a cold start and from closed-form warm starts computed from the pair correlations -- which are the
negative gradient at x = 0, i.e. a by-product of the solver's first pass.  Result (N = 48 / 96, 4-regular J = +-0.4,
1e5 Gibbs samples, tol 1e-6): cold 31 / 42 rounds; naive mean-field start (J = -C^-1) 22 / 28 rounds; the
independent-pair start does not help.  See DESIGN.md section 8.     usage: python scripts/dev_warmstart_study.py [N] [K]"""
import pathlib, sys
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from bench import regular_graph
rng=np.random.default_rng(0)
N=int(sys.argv[1]) if len(sys.argv)>1 else 48
K=int(float(sys.argv[2])) if len(sys.argv)>2 else 200000
lo,hi=regular_graph(N,4,N)
J=np.zeros((N,N)); sg=rng.choice([-1.,1.],size=len(lo))
for i,j,s in zip(lo,hi,sg): J[i,j]=J[j,i]=0.4*s
# Gibbs sampling, many chains
S=rng.choice([-1,1],size=(K,N)).astype(np.int8)
for sweep in range(40):
    for i in range(N):
        fld=S@J[i]
        S[:,i]=np.where(rng.random(K)<1/(1+np.exp(-2*fld)),1,-1)
Sf=S.astype(np.float64)
w=np.full(K,1.0/K)
lam=0.4*np.sqrt(np.log(N*N/0.05)/K)
Q=np.concatenate([Sf,np.ones((K,1))],axis=1)   # features [S|1]
F=N+1
pen=np.ones((N,F)); pen[:,N]=0            # L1 on couplings, field free
fixed=np.zeros((N,F),bool); fixed[np.arange(N),np.arange(N)]=True
def fg(X):   # X [N,F] -> f[N], G[N,F]
    E=Q@X.T            # [K,N]
    T=Sf*E
    psi=np.exp(-T)*w[:,None]
    f=psi.sum(0)
    G=-(Sf*psi).T@Q
    G[fixed]=0
    return f,G
def fista(X0,tol=1e-6,max_it=400):
    X=X0.copy(); Y=X0.copy(); L=np.ones(N); t=np.ones(N); active=np.ones(N,bool)
    fY,G=fg(Y); rounds=0; streak=np.zeros(N,int)
    Z=X.copy()
    hist=[]
    while rounds<max_it:
        thr=lam/L
        Zc=Y-G/L[:,None]
        Zs=np.sign(Zc)*np.maximum(np.abs(Zc)-thr[:,None],0)
        Zn=np.where(pen>0,Zs,Zc); Zn[fixed]=0
        gm=L*np.abs(Zn-Y).max(1)
        r=((Y-Zn)*(Zn-X)).sum(1)
        restart=r>0
        tn=np.where(restart,1.0,0.5*(1+np.sqrt(1+4*t*t)))
        beta=np.where(restart,0.0,(t-1)/tn)
        Yn=Zn+beta[:,None]*(Zn-X)
        conv=gm<=tol
        newly=active&conv
        X[newly]=Zn[newly]; active&=~conv
        hist.append(active.sum())
        if not active.any(): break
        fN,GN=fg(Yn); rounds+=1
        d=Yn-Y
        q1=(G*d).sum(1); c=0.5*L*(d*d).sum(1)
        D=fY+q1+c-fN
        reject=active&(D< -0.1*c)
        acc=active&~reject
        L[reject]*=2; streak[reject]=0
        X[acc]=Zn[acc]; Y[acc]=Yn[acc]; G[acc]=GN[acc]; fY[acc]=fN[acc]; t[acc]=tn[acc]
        good=acc&(c>0)&(D>0.25*c)
        streak=np.where(good,streak+1,np.where(acc,0,streak))
        relax=streak>=3; L[relax]*=0.85; streak[relax]=0
    return X,rounds,hist
X0=np.zeros((N,F))
Xc,rc,hc=fista(X0)
print("cold rounds",rc, "active trace", hc[::4])
# independent-pair warm start
m=Sf.mean(0); C=(Sf.T@Sf)/K
with np.errstate(all='ignore'):
    num=(1+m[:,None]+m[None,:]+C)*(1-m[:,None]-m[None,:]+C)
    den=(1+m[:,None]-m[None,:]-C)*(1-m[:,None]+m[None,:]-C)
    Jip=0.25*np.log(num/den)
np.fill_diagonal(Jip,0)
for thr in (0.0,0.05,0.1,0.2):
    Xw=np.zeros((N,F)); Xw[:,:N]=np.where(np.abs(Jip)>thr,Jip,0)
    Xs,rw,hw=fista(Xw)
    print(f"warm IP thr {thr}: rounds {rw}, |x0-x*|max {np.abs(Xw-Xc).max():.3f}, diff sol {np.abs(Xs-Xc).max():.2e}")
# naive mean-field
Cc=C-np.outer(m,m)
Jmf=-np.linalg.inv(Cc); np.fill_diagonal(Jmf,0)
for thr in (0.0,0.1,0.2):
    Xw=np.zeros((N,F)); Xw[:,:N]=np.where(np.abs(Jmf)>thr,Jmf,0)
    Xs,rw,hw=fista(Xw)
    print(f"warm nMF thr {thr}: rounds {rw}, |x0-x*|max {np.abs(Xw-Xc).max():.3f}")

# Sessak-Monasson small-correlation expansion (nMF + independent pair - loop term)
with np.errstate(all='ignore'):
    Lm = 1 - m * m
    Jsm = Jmf + Jip - Cc / (np.outer(Lm, Lm) - Cc * Cc)
np.fill_diagonal(Jsm, 0)
Xw = np.zeros((N, F)); Xw[:, :N] = Jsm
Xs, rw, hw = fista(Xw)
print(f"warm SM: rounds {rw}, |x0-x*|max {np.abs(Xw-Xc).max():.3f}")
