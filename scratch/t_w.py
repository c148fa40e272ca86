import sys; sys.path.insert(0,"/root/repo")
import numpy as np
from oracle import weighting as ow
g = np.load("tests/golden/weighting.npz")
uvw,freq,mask = g["uvw"],g["freq"],g["mask"]; nx,ny,cell=int(g["nx"]),int(g["ny"]),float(g["cell"])
for tag,dt in (("f8",np.float64),("f4",np.float32)):
    wgt=g[f"wgt_{tag}"]
    for us,vs in ((-1.0,1.0),(1.0,-1.0)):
        st=f"{tag}_{int(us)}_{int(vs)}"
        c = ow.compute_counts(uvw,freq,mask,wgt,nx,ny,cell,cell,dt,1,us,vs)
        print(st,"counts maxdiff",np.abs(c-g[f"counts_{st}"]).max(), "nonzero pattern equal", np.array_equal(c>0, g[f"counts_{st}"]>0))
        for r in (-2.0,0.0,1.5):
            c2=g[f"counts_{st}"].copy(); w2=wgt.copy()
            ow.counts_to_weights(c2,uvw,freq,w2,mask,nx,ny,cell,cell,r,us,vs)
            print("   r",r,"w rel",np.abs(w2-g[f"w_{st}_r{r}"]).max()/np.abs(w2).max(),"c rel",np.abs(c2-g[f"c_{st}_r{r}"]).max()/np.abs(c2).max())
    print("filter",np.abs(ow.filter_extreme_counts(g[f"counts_{tag}_-1_1"].copy(),10.0)-g[f"filtered_{tag}"]).max(),
          "box",np.abs(ow.box_sum_counts(g[f"counts_{tag}_-1_1"].copy(),2)-g[f"boxsum_{tag}"]).max())
