"""Debug (torchrun, 2 GPUs): first field / step where the sharded filter and the one-GPU filter differ bitwise."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine, ShardedParticleFilterEngine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
outs = ("propagated", "norm_weights", "ancestors", "particles", "means", "log_likelihood", "ess")
scal = ("means", "log_likelihood", "ess")
c = case("bot2"); T, N = 8, 8192 * world
_, ys = c.simulate(T, 7); u = c.inputs(T)
x0, noise, ru = c.pf_randomness(T, N, 8, "systematic")
Ns = N // world; sl = slice(rank * Ns, (rank + 1) * Ns)
eng = ShardedParticleFilterEngine(c.host_model(), N, 1.0, "systematic", seed=3, device=local)
out = eng.run_host(ys, inputs=u, init_particles=x0[sl], noise=noise[:, sl], resample_u=ru, outputs=outs)
eng.close()
full = {}
for k, v in out.items():
    if k in scal:
        t = torch.from_numpy(np.ascontiguousarray(v)).cuda(); dist.broadcast(t, src=0); full[k] = t.cpu().numpy()
    else:
        parts = [None] * world; dist.all_gather_object(parts, v); full[k] = np.concatenate(parts, axis=2)
if rank == 0:
    one = ParticleFilterEngine(c.host_model(), N, 1, 1.0, "systematic", "streamed", 3, local)
    o1 = one.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=outs)
    one.close()
    for t in range(T):
        line = [f"t={t}"]
        for k in outs:
            a, b = full[k][0, t], o1[k][0, t]
            nd = int(np.sum(a != b))
            md = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if nd else 0.0
            line.append(f"{k}:{nd}/{a.size} max {md:.3g}")
        print("  ".join(line), flush=True)
dist.barrier(); dist.destroy_process_group()
