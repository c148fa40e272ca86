import numpy as np, time
from numpy.polynomial.legendre import leggauss

def phi(x, beta):
    return np.exp(beta*(np.sqrt(np.maximum(1-x*x,0))-1))

_nodes, _wts = leggauss(128)
def psihat(xi, W, beta):
    # (W/2) int_{-1}^{1} phi(x) cos(pi W xi x) dx
    x = _nodes; w = _wts
    return (W/2)*np.sum(w[None,:]*phi(x,beta)[None,:]*np.cos(np.pi*W*np.outer(xi,x)),axis=1)

def kerr(W, beta, sigma, nxi=48, ng=24):
    xi = np.linspace(0, 0.5/sigma, nxi)
    g = (np.arange(ng)+0.5)/ng  # frac offsets
    # taps j = ceil(g - W/2) ... : j0 = floor(g - W/2)+1
    j0 = np.floor(g - W/2).astype(int)+1
    j = j0[:,None]+np.arange(W)[None,:]       # (ng, W)
    t = g[:,None]-j                            # in [-W/2, W/2)
    psi = phi(2*t/W, beta)                     # (ng,W)
    ph = np.exp(2j*np.pi*j[None,:,:]*xi[:,None,None])  # (nxi,ng,W)
    approx = np.sum(psi[None]*ph,axis=2)/psihat(xi,W,beta)[:,None]
    exact = np.exp(2j*np.pi*g[None,:]*xi[:,None])
    e = np.abs(approx-exact)
    return e.max(), np.sqrt((e**2).mean(axis=1)).max()

def best(W, sigma):
    b0 = np.pi*W*(1-1/(2*sigma))
    best=(1e9,None)
    for gam in np.linspace(0.85,1.05,41):
        e = kerr(W, gam*b0, sigma)[0]
        if e<best[0]: best=(e,gam)
    return best
t=time.time()
for sigma in (1.25,1.5,2.0):
    for W in range(4,17):
        e,gam = best(W,sigma)
        print(sigma,W,"%.3e"%e,"gam=%.3f beta/W=%.3f"%(gam, gam*np.pi*(1-1/(2*sigma))))
print(time.time()-t)
print("rms metric")
def best2(W, sigma):
    b0 = np.pi*W*(1-1/(2*sigma))
    best=(1e9,None)
    for gam in np.linspace(0.85,1.05,81):
        e = kerr(W, gam*b0, sigma)[1]
        if e<best[0]: best=(e,gam)
    return best
for sigma in (1.2, 1.3,1.5,1.75,2.0):
    print(sigma, ["%d:%.1e"%(W,best2(W,sigma)[0]) for W in range(4,17)])
