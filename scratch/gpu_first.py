import sys, itertools, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
from pfb_imaging_b200 import wgridder as W, synth
from pfb_imaging_b200.plan import make_plan, w_range
from oracle import dft, wgridder_np as wg

def kat():
    np.random.seed(42)
    npix=1024; num_ants=100
    pixsize = 0.5*np.pi/180/3600
    a1,a2 = np.asarray(list(itertools.combinations(range(num_ants),2))).T
    ants = 10e3*np.random.normal(size=(num_ants,3)); ants[:,2]*=0.001
    uvw = ants[a1]-ants[a2]
    freqs = np.linspace(700e6,2000e6,2)
    dirty=np.zeros((npix,npix)); dirty[npix//2,npix//2]=1; dirty[npix//4,npix//4]=1
    for prec,eps in (("double",1e-6),("single",1e-5),("double",1e-10)):
      for (l0,m0) in [(0,0),(0.1,-0.17),(0.2,0.5)]:
        flips=(False,True,False)
        fu,fv,fw=flips
        x = dirty.astype(np.float32 if prec=="single" else np.float64)
        t=time.time()
        gp = W.plan_for(uvw,freqs,npix_x=npix,npix_y=npix,pixsize_x=pixsize,pixsize_y=pixsize,center_x=-l0,center_y=-m0,
                epsilon=eps,flip_u=fu,flip_v=fv,flip_w=fw,do_wgridding=True,divide_by_n=True,precision=prec)
        v = gp.degrid(x)
        t1=time.time()-t
        ref = dft.dft_dirty2vis(uvw,freqs,dirty,pixsize,pixsize,-l0,-m0,fu,fv,fw,True,True)
        err = np.linalg.norm(v-ref)/np.linalg.norm(ref)
        p=gp.plan
        print(prec,eps,(l0,m0),"W",p.W,"sig",p.sigma,"nu",p.nu,"P",p.nplanes,"degrid relL2 %.2e maxabs %.2e"%(err,np.abs(v-ref).max()),"%.2fs"%t1, flush=True)
        rng=np.random.default_rng(1)
        vis = (rng.standard_normal(v.shape)+1j*rng.standard_normal(v.shape)).astype(v.dtype)
        wgt = rng.uniform(0.5,1.5,v.shape).astype(x.dtype)
        d = gp.grid(vis,wgt)
        px = (rng.integers(0,npix,50), rng.integers(0,npix,50))
        dref = dft.dft_vis2dirty(uvw,freqs,vis,wgt,None,npix,npix,pixsize,pixsize,-l0,-m0,fu,fv,fw,True,True,pixels=px)
        adj = abs(np.vdot(v.astype(np.complex128), (vis*wgt).astype(np.complex128)).real - np.sum(d.astype(np.float64)*dirty))/abs(np.sum(d.astype(np.float64)*dirty))
        print("   grid relL2 %.2e"%(np.linalg.norm(d[px]-dref)/np.linalg.norm(dref)), "adjoint %.2e"%adj, flush=True)
        # hessian vs degrid+grid
        gp.bind_weights(wgt)
        h = gp.hessian(x)
        h2 = gp.grid(gp.degrid(x), wgt)
        print("   hessian vs composed %.2e"%(np.linalg.norm(h-h2)/np.linalg.norm(h2)))
        if prec=="double" and eps==1e-6:
            b = wg.bin_indices(p, uvw, freqs)
            dmp = gp.bin_dump()
            ok = all(np.array_equal(dmp[k], b[k]) for k in ("iu0","iv0","ip0","key"))
            ok2 = np.array_equal(dmp["sorted_idx"].astype(np.int64), b["idx"][b["order"]])
            print("   bin bit-exact:", ok, ok2)
        gp.close()

def perf(nx, ntime, nchan, prec, eps):
    import ctypes
    d = synth.make_band(ntime, nchan, band=7, precision=prec, with_vis=False)
    uvw, freq = d["uvw"], d["freq"]
    cell = synth.default_cell(uvw, 1712e6)
    print("cell", cell, "nrow", uvw.shape[0], "nvis", uvw.shape[0]*nchan, flush=True)
    t=time.time()
    gp = W.plan_for(uvw,freq,npix_x=nx,npix_y=nx,pixsize_x=cell,pixsize_y=cell,epsilon=eps,flip_v=True,
          do_wgridding=True,divide_by_n=False,precision=prec,sigma_min=1.1,sigma_max=3.0)
    print("plan+bind %.2fs"%(time.time()-t), gp.info(), flush=True)
    gp.bind_weights(d["wgt"])
    x = synth.point_source_image(nx,nx,dtype=gp.rdt)
    gp.set_profiling(True)
    for it in range(3):
        t=time.time()
        h = gp.hessian(x)
        print("hessian %.1f ms"%(1e3*(time.time()-t)), ["%.2f"%m for m in gp.timings()], flush=True)
    # check against sampled DFT: degrid
    v = gp.degrid(x)
    rows = np.random.default_rng(0).integers(0, uvw.shape[0], 200)
    ref = dft.dft_dirty2vis(uvw,freq,x.astype(np.float64),cell,cell,0,0,False,True,False,True,False,rows=rows)
    print("degrid relL2 vs DFT %.2e"%(np.linalg.norm(v[rows]-ref)/np.linalg.norm(ref)), flush=True)
    print("degrid timings", ["%.2f"%m for m in gp.timings()])
    gp.close()

if __name__=="__main__":
    what = sys.argv[1] if len(sys.argv)>1 else "all"
    if what in ("kat","all"): kat()
    if what in ("perf","all"):
        perf(2048, 62, 8, "double", 1e-5)
        perf(4096, 775, 16, "single", 1e-5)
