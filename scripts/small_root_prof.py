import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
os.environ["PC_SMALL_PROF"] = "1"
r = bench.time_small_block_batch(torch.device("cuda", 0), reps=2)
print({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a in ("ms", "newton_iters_mean")}) for k, v in r.items()})
