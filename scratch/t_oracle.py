import sys, itertools, time
sys.path.insert(0, "/root/repo")
import numpy as np
from pfb_imaging_b200.plan import make_plan, w_range
from oracle import dft, wgridder_np as wg

np.random.seed(42)
npix=1024; num_ants=100
pixsize = 0.5*np.pi/180/3600
a1,a2 = np.asarray(list(itertools.combinations(range(num_ants),2))).T
ants = 10e3*np.random.normal(size=(num_ants,3)); ants[:,2]*=0.001
uvw = ants[a1]-ants[a2]
freqs = np.linspace(700e6,2000e6,2)
dirty=np.zeros((npix,npix)); dirty[npix//2,npix//2]=1; dirty[npix//4,npix//4]=1
uvw = uvw[::10]
for (l0,m0) in [(0,0),(0.1,-0.17),(0.2,0.5)]:
  for flips in [(False,False,False),(False,True,False)]:
    fu,fv,fw = flips
    eps=1e-6
    wmin,wmax = w_range(uvw,freqs,-1.0 if fw else 1.0)
    t=time.time()
    plan = make_plan(nx=npix,ny=npix,pixsize_x=pixsize,pixsize_y=pixsize,center_x=-l0,center_y=-m0,epsilon=eps,
        flip_u=fu,flip_v=fv,flip_w=fw,do_wgridding=True,divide_by_n=True,wmin=wmin,wmax=wmax,nvis=uvw.shape[0]*2, sigma_max=1.5)
    ref = dft.dft_dirty2vis(uvw,freqs,dirty,pixsize,pixsize,-l0,-m0,fu,fv,fw,True,True)
    v = wg.dirty2vis_np(plan,uvw,freqs,dirty)
    err = np.linalg.norm(v-ref)/np.linalg.norm(ref)
    print((l0,m0),flips,"W",plan.W,"sig",plan.sigma,"nu",plan.nu,"P",plan.nplanes,"degrid relL2 %.2e maxabs %.2e"%(err,np.abs(v-ref).max()), "%.1fs"%(time.time()-t))
    # adjoint
    rng=np.random.default_rng(1)
    vis = rng.standard_normal(v.shape)+1j*rng.standard_normal(v.shape)
    wgt = rng.uniform(0.5,1.5,v.shape)
    d = wg.vis2dirty_np(plan,uvw,freqs,vis,wgt)
    px = (rng.integers(0,npix,50), rng.integers(0,npix,50))
    dref = dft.dft_vis2dirty(uvw,freqs,vis,wgt,None,npix,npix,pixsize,pixsize,-l0,-m0,fu,fv,fw,True,True,pixels=px)
    print("   grid relL2 %.2e"%(np.linalg.norm(d[px]-dref)/np.linalg.norm(dref)),
          "adjoint %.2e"%(abs(np.vdot(v, vis*wgt) - np.sum(d*dirty))/abs(np.sum(d*dirty))))
