"""The reference's own resampler (multinomial, `jr.choice`) at the bench size: BOT, N = 2^log2N, Philox noise and uniforms.
Positions are unsorted, so every slot bisects tau in global memory (25 dependent probes at N = 2^24).
usage: python profiles/run_pf_multinomial.py [T] [log2N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
N = 1 << log2n
c = case("bot2")
_, ys = c.simulate(T, 1000)
u = c.inputs(T)
dev = torch.device("cuda", 0)
ys_d, u_d = torch.from_numpy(ys).to(dev), torch.from_numpy(u).to(dev)
for res in ("systematic", "multinomial"):
    eng = ParticleFilterEngine(c.host_model(), N, 1, 1.0, res, "streamed", 1234, 0)
    for _ in range(2):
        eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.phase_times(reset=True)
    e0.record()
    for _ in range(3):
        out = eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ph = {k: round(v * 1e3 / (3 * T), 1) for k, v in eng.phase_times().items()}
    print(f"{res:12s} N=2^{log2n} T={T}: {ms / T * 1e3:.1f} us per step, {N * T / (ms * 1e-3):.3g} particle-steps/s, phases (us/step) {ph}", flush=True)
    eng.close()
