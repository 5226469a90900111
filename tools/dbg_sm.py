import sys; sys.path.insert(0, '.')
import numpy as np
import uni_b200 as u
from oracle import mat_oracle as mo
u.init(0)
T,N=5000,40
rng=np.random.default_rng(5); a=rng.standard_normal((T,N))
a[17,1]=np.nan;a[0,2]=np.nan;a[4999,3]=np.inf;a[0,4]=np.inf;a[123,5]=-np.inf;a[0,6]=-np.inf;a[:,7]=-np.inf;a[2500,8]=900.0;a[4000,9]=-900.0
a[:,10]=np.linspace(-400,400,T);a[0,17]=1e308;a[1,17]=-1e308;a[4999,20]=350.0;a[:,21]=0.0;a[3,33]=-np.inf;a[4998,33]=np.nan
w=mo.softmax(mo.from2d(a),0).to2d()
for path in (1,2):
    u.set_softmax_axis0_path(path)
    g=u.Mat.from_numpy(a).softmax(0).to_numpy()
    bad=np.where((np.isnan(g)!=np.isnan(w)).any(axis=0))[0]
    print("path",path,"nan-mismatch lanes",bad)
    for j in bad: print(j, g[:3,j], w[:3,j], np.isnan(g[:,j]).sum())
    ok=~np.isnan(w)
    rel=np.abs(g[ok]-w[ok])/np.maximum(np.abs(w[ok]),1e-300)
    print(" max rel err", np.nanmax(rel))
