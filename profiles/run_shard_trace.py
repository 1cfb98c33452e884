"""ONE particle filter sharded over the ranks (torchrun): per-rank phase times.  usage: torchrun ... profiles/run_shard.py [T] [log2N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bayesianfiltering_b200 import workloads
from bayesianfiltering_b200.engine import ShardedParticleFilterEngine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = workloads.bot()
_, ys = wl.sample(T, 1000)
u = wl.inputs(T)
ys_d, u_d = torch.from_numpy(ys).to(dev), torch.from_numpy(u).to(dev)
eng = ShardedParticleFilterEngine(wl.host_model(), N, 1.0, "systematic", seed=1234, device=local)
for _ in range(2):
    out = eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
torch.cuda.synchronize(); dist.barrier()
import os as _os
_os.environ["BF_TRACE_FILE"] = _os.environ.get("BF_TRACE_FILE", "gpurun_out/trace") + f"_r{rank}.bin"
eng.phase_times(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    out = eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
e1.record(); torch.cuda.synchronize(); dist.barrier()
eng.lib.bf_pf_phase_times; ph = {k: round(v * 1e3 / (2 * T), 1) for k, v in eng.phase_times().items()}
print(f"rank {rank}/{world} us/step {e0.elapsed_time(e1) * 1e3 / (2 * T):.1f} {ph} env STAGE={os.environ.get('BF_SHARD_STAGE')} MODE={os.environ.get('BF_SHARD_MODE')}", flush=True)
import subprocess
subprocess.run([sys.executable, "profiles/trace_summary.py", _os.environ["BF_TRACE_FILE"]])
eng.close()
dist.destroy_process_group()
