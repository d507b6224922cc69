import sys, torch
sys.path.insert(0, ".")
import bench
r = bench.time_small_block_batch(torch.device("cuda", 0), reps=2)
print(r["persistent_tcgen05"]["ms"])
