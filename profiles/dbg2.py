import sys; sys.path.insert(0,"/root/repo")
import numpy as np
from tests.parity_cases import case
from oracle import pf as opf
from bayesianfiltering_b200.engine import ParticleFilterEngine
c=case("ungm"); T,N=8,1000
_,ys=c.simulate(T,7); u=c.inputs(T); x0,noise,ru=c.pf_randomness(T,N,8,"multinomial")
outs=("weights","particles","propagated","norm_weights","ancestors","resampled","ess","log_likelihood")
o={}
for path in ("resident","streamed"):
    e=ParticleFilterEngine(c.host_model(),N,1,0.5,"multinomial",path,0,0)
    o[path]=e.run_host(ys,inputs=u,init_particles=x0,noise=noise,resample_u=ru,outputs=outs); e.close()
print("resampled", o["resident"]["resampled"][0], o["streamed"]["resampled"][0])
for t in range(T):
    a=o["resident"]["norm_weights"][0,t]; b=o["streamed"]["norm_weights"][0,t]
    d=np.abs(a-b)/np.maximum(a,1e-30)
    print(t, "max rel diff nw", d.max(), "sum", b.sum(), "ll diff", o["resident"]["log_likelihood"][0,t]-o["streamed"]["log_likelihood"][0,t], "prop diff", np.abs(o["resident"]["propagated"][0,t]-o["streamed"]["propagated"][0,t]).max())
    if d.max()>1e-3:
        i=np.argmax(d); print("   idx",i,a[i],b[i], "ratio", (b/np.maximum(a,1e-30))[:8]); break
t=3
a=o["resident"]["norm_weights"][0,t]; b=o["streamed"]["norm_weights"][0,t]
print("sorted equal?", np.allclose(np.sort(a), np.sort(b), rtol=1e-5), "nonzero counts", (a>0).sum(), (b>0).sum())
# does streamed nw[t] equal resident nw of some other step, or a blockwise permutation?
for blk in (4, 32, 128):
    A=a.reshape(-1)[: (N//blk)*blk].reshape(-1,blk); B=b.reshape(-1)[: (N//blk)*blk].reshape(-1,blk)
    match=[(np.abs(A - B[i]).max(axis=1) < 1e-9).nonzero()[0] for i in range(min(8, len(B)))]
    print(blk, match)
w2=o["streamed"]["weights"][0,2]; print("weights[2] uniform?", np.all(w2==w2[0]))
# is it exp(ll-mx)*W_garbage ?  ratio b/a vs streamed norm_weights of earlier steps
r=b/np.maximum(a,1e-30)
for tt in range(3):
    pw=o["streamed"]["norm_weights"][0,tt]
    print("ratio ~ N*nw[",tt,"]?", np.allclose(r/ r.mean(), pw/pw.mean(), rtol=1e-3))
