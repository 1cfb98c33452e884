"""Run under torchrun on >= 2 GPUs (tests/test_gpu_shard.py launches it): ONE particle filter sharded over the ranks
against the oracle, step by step, with the randomness supplied -- the same checks as tests/test_gpu_pf._stepwise,
on arrays gathered from the shards.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)



def _same_filter(sharded, one, what):
    """The sharded filter IS the one-GPU filter: same bits, whole run.  Philox counters, resampling positions and ancestors
    are indexed by the global particle; the particle-filter translation units are compiled with -fmad=false (every fused
    multiply-add is spelled out), so the sharded and the one-GPU instantiation round alike; and the step statistics (max,
    normaliser, ESS, weighted mean) are reduced over a canonical binary tree of the particle index whose aligned subtrees are
    the shards (pf_fused.cuh, PfFusedArgs::upart), as the cdf already was."""
    for k in ("log_likelihood", "means", "resampled"):
        if k in sharded and k in one:
            assert np.array_equal(np.asarray(sharded[k]), np.asarray(one[k])), f"{what}: {k} of the sharded filter differs from the one-GPU filter"


def main():
    import torch
    import torch.distributed as dist
    from tests.parity_cases import case, close
    from tests.test_gpu_pf import _stepwise
    from bayesianfiltering_b200.engine import ShardedParticleFilterEngine, ParticleFilterEngine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    outs = ("weights", "particles", "propagated", "norm_weights", "ancestors", "means", "log_likelihood", "ess",
            "resampled", "final_particles", "final_weights")
    scal = ("means", "log_likelihood", "ess", "resampled")
    for name, T, N, thr, res in (("lg42", 10, 4096 * world, 0.6, "systematic"), ("bot2", 8, 8192 * world, 1.0, "systematic"),
                                 ("ungm", 10, 2048 * world, 0.5, "multinomial")):
        c = case(name)
        _, ys = c.simulate(T, 7)
        u = c.inputs(T)
        x0, noise, ru = c.pf_randomness(T, N, 8, res)               # identical on every rank (seeded)
        Ns = N // world
        sl = slice(rank * Ns, (rank + 1) * Ns)
        eng = ShardedParticleFilterEngine(c.host_model(), N, thr, res, seed=3, device=local)
        out = eng.run_host(ys, inputs=u, init_particles=x0[sl], noise=noise[:, sl],
                           resample_u=(ru[:, sl] if res == "multinomial" else ru), outputs=outs)
        eng.close()
        # assemble the whole filter on every rank
        full = {}
        for k, v in out.items():
            if k in scal:
                t = torch.from_numpy(np.ascontiguousarray(v)).cuda()
                dist.broadcast(t, src=0)
                full[k] = t.cpu().numpy()
            else:
                parts = [None] * world
                dist.all_gather_object(parts, v)
                full[k] = np.concatenate(parts, axis=1 if k.startswith("final_") else 2)      # the particle axis
        if rank == 0:
            n_res = _stepwise(c, full, ys, u, x0, noise, ru, thr, res)
            assert n_res > 0, "case never resampled"
            # and against the unsharded filter on one GPU: same decisions, same estimates
            one = ParticleFilterEngine(c.host_model(), N, 1, thr, res, "streamed", 3, local)
            o1 = one.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru,
                              outputs=("means", "log_likelihood", "resampled", "ancestors", "final_particles"))
            one.close()
            _same_filter(full, o1, name)
            assert np.array_equal(full["ancestors"], o1["ancestors"]) and np.array_equal(full["final_particles"], o1["final_particles"])
            print(f"sharded {name}: N={N} over {world} ranks, T={T}, {n_res} resampling steps: ok", flush=True)
        dist.barrier()
    # Philox mode: the random numbers are indexed by GLOBAL particle, so the filter does not depend on the sharding
    # (2^21 per rank: 1024 tiles over 147 CTAs -- every CTA owns several blocks dealt round the grid, pf_fused.cuh "Ownership";
    #  summaries only: the lean sharded instantiation; with ancestors: the general one)
    c = case("bot2")
    for T, N, outs in ((40, 16384 * world, ("means", "log_likelihood")), (12, (1 << 21) * world, ("means", "log_likelihood")),
                       (6, (1 << 21) * world, ("means", "log_likelihood", "resampled", "ancestors"))):
        _, ys = c.simulate(T, 9)
        u = c.inputs(T)
        eng = ShardedParticleFilterEngine(c.host_model(), N, 1.0, "systematic", seed=11, device=local)
        out = eng.run_host(ys, inputs=u, outputs=outs)
        eng.close()
        if rank == 0:
            one = ParticleFilterEngine(c.host_model(), N, 1, 1.0, "systematic", "streamed", 11, local)
            o1 = one.run_host(ys, inputs=u, outputs=outs)
            one.close()
            for k in outs:
                ref = np.asarray(o1[k])
                if k == "ancestors": ref = ref[:, :, :N // world]          # rank 0 holds its own slots (global ancestor indices)
                assert np.array_equal(np.asarray(out[k]), ref), f"philox N={N}: {k} differs from the one-GPU filter"
            print(f"sharded philox run N={N} equals the one-GPU run: ok", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
