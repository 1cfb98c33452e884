"""What a random 16-byte gather costs on this GPU (the floor under multinomial resampling, whose ancestors are unsorted):
rows of 4 floats taken from a [2^24, 4] array at 2^24 random indices vs at sorted indices (systematic resampling).
Library call (torch.index_select), used only as a yardstick; nothing in the product calls it."""
import torch
dev = torch.device("cuda", 0)
N = 1 << 24
x = torch.randn(N, 4, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
idx_r = torch.randint(0, N, (N,), device=dev, generator=g, dtype=torch.int64)
idx_s = torch.sort(idx_r).values
for name, idx in (("sorted", idx_s), ("random", idx_r)):
    for _ in range(3): torch.index_select(x, 0, idx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.index_select(x, 0, idx)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{name:7s} gather of 2^24 x 16 B: {us:7.1f} us  ({N * 40 / us / 1e3:.0f} GB/s counting 8 B index + 16 B read + 16 B write)")
