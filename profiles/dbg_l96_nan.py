"""Which trials of the Lorenz-96 Gaussian-sum filter go non-finite, on package-generated data, d=40 kernel vs generic kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianfiltering_b200 import workloads
from bayesianfiltering_b200.engine import GaussianSumFilterEngine
M, B, T = 256, 512, 100
l = workloads.l96(40)
xs, ys = l.sample(T, 700, num_trials=B)
print("data finite", np.isfinite(xs).all(), np.isfinite(ys).all(), "state range", xs.min(), xs.max())
m0 = l.init_means(M, B, 800)
dev = torch.device("cuda", 0)
for generic in ("0", "1"):
    os.environ["BF_GSF_LARGE_GENERIC"] = generic
    eng = GaussianSumFilterEngine(l.host_model(), M, B, 0)
    out = eng.run(torch.from_numpy(ys).to(dev), init_means=torch.from_numpy(m0).to(dev), outputs=("log_likelihood", "mixture_mean"))
    torch.cuda.synchronize()
    ll = out["log_likelihood"].cpu().numpy()
    bad = ~np.isfinite(ll)
    first = np.where(bad.any(axis=1), bad.argmax(axis=1), -1)
    print("generic" if generic == "1" else "d40", "non-finite trials:", int(bad.any(axis=1).sum()), "first bad steps", sorted(set(first[first >= 0]))[:10],
          "loglik of finite trials", float(ll[~bad.any(axis=1)].sum()))
    mm = out["mixture_mean"].cpu().numpy()
    ok = ~bad.any(axis=1)
    print("   rmse of the mixture mean vs truth over finite trials", float(np.sqrt(np.mean((mm[ok] - xs[ok]) ** 2))))
    eng.close()
