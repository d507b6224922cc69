#!/usr/bin/env python
"""Headline benchmark: batched inverse 4th roots of 1024x1024 Shampoo statistics.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

A "step" is one pass of the preconditioner-root hot path
(`_matrix_inverse_pth_root_vmap`, DS:2742-2744) over one batch of synthetic SPD
statistics.  With N GPUs (torchrun, one rank per GPU) every rank owns its own
batch (weak scaling, blocks partitioned across ranks as DS:2862-2875 does) and
the step ends with the all-gather of the roots over NCCL (DS:2876).

Prints ONE JSON line (rank 0).  See the task contract for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "inverse_pth_roots_per_sec"
UNIT = "roots/s"


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--n", type=int, default=1024, help="statistic size (block_size)")
  ap.add_argument("--batch", type=int, default=74,
                  help="statistics per GPU per step (74 = one per SM pair of a B200)")
  ap.add_argument("--p", type=int, default=4)
  ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc6", "tc3", "fp16x3"])
  ap.add_argument("--cpu-sample", type=int, default=2,
                  help="matrices in the bounded CPU-baseline sample")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-step", action="store_true", help="skip the Shampoo-step measurement")
  ap.add_argument("--no-big", action="store_true",
                  help="skip the BERT-large (config 4) and Sketchy-step (config 5) measurements")
  ap.add_argument("--split", type=int, default=0,
                  help="sub-batches per step (default 1; with k > 1 the all-gather of sub-batch i "
                       "overlaps the solve of i+1, measured slower: each sub-batch pays its own "
                       "power iteration and launch ramps)")
  return ap.parse_args()


def workload_name(a):
  return f"ema_lowrank_statistics_{a.n}x{a.n}_p{a.p}_batch{a.batch}_per_gpu"


def workload_config(a, world):
  """`config` of the JSON line -- identical keys and values in both arms (`--impl ours` /
  `--impl reference`); run-specific facts (engine, iteration counts, ...) go to `run_info`."""
  return {"workload": workload_name(a), "n": a.n, "p": a.p, "batch_per_gpu": a.batch,
          "statistics_class": "S0=1e-6 I; 20 EMA steps of rank n/32 Gram updates (cond ~1e6)",
          "l2": "inputs_exceed_l2" if a.batch * a.n * a.n * 4 > 126e6 else "inputs_fit_l2",
          "sharding": (f"blocks partitioned over {world} ranks + all_gather" if world > 1
                       else "single gpu")}


def newton_gemms(p):
  """G(p), SURVEY 8(d)."""
  return (int(p).bit_length() - 1) + bin(int(p)).count("1") - 1 + 2


def make_statistics_torch(batch, n, seed, device):
  """Synthetic SPD statistics: S0 = 1e-6 I; S <- 0.999 S + 0.001 G G^T, 20 steps of
  G ~ N(0, 1) [n x n/32] (SURVEY 8(d) class (ii); DS:2594, DS:2635-2636).  The
  accumulated Gram has rank 20 n/32 < n, so the spectrum has a 1e-6 floor and a
  condition number ~1e6 like real early-training Shampoo statistics; the coupled
  Newton iteration needs ~19 steps on it (vs ~6 for a well-conditioned Wishart)."""
  import torch
  gen = torch.Generator(device=device)
  gen.manual_seed(seed)
  out = torch.empty((batch, n, n), dtype=torch.float32, device=device)
  eye = torch.eye(n, device=device, dtype=torch.float64)
  for b in range(batch):
    s = 1e-6 * eye
    for _ in range(20):
      g = torch.randn((n, max(n // 32, 1)), generator=gen, device=device,
                      dtype=torch.float32).double()
      s = 0.999 * s + 0.001 * (g @ g.T)
    out[b] = s.float()
  return out


def make_statistics_numpy(batch, n, seed):
  rng = np.random.default_rng(seed)
  out = np.empty((batch, n, n), np.float32)
  for b in range(batch):
    s = 1e-6 * np.eye(n)
    for _ in range(20):
      g = rng.standard_normal((n, max(n // 32, 1))).astype(np.float32).astype(np.float64)
      s = 0.999 * s + 0.001 * (g @ g.T)
    out[b] = s.astype(np.float32)
  return out


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
  FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
            "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.rows, self.proc = [], None
    self.gpu = gpu_index

  def start(self):
    # NVML in a thread (one query ~0.1 ms, polled every 5 ms) so that even a 100 ms timed
    # region is sampled tens of times; nvidia-smi (slow to start) is the fallback.
    self.nvml_rows, self.stop_flag, self.thread = [], False, None
    try:
      import pynvml
      pynvml.nvmlInit()
      h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(pynvml))
      reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
          pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
      bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20,
              "hw_thermal_slowdown": 0x40}

      def poll():
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not self.stop_flag:
          try:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            mask = int(reasons_fn(h))
            self.nvml_rows.append((float(sm), float(mx), [k for k, b in bits.items() if mask & b]))
          except pynvml.NVMLError:
            pass
          time.sleep(0.005)

      self.thread = threading.Thread(target=poll, daemon=True)
      self.thread.start()
      self.proc = None
      return
    except Exception:  # pylint: disable=broad-except
      self.thread = None
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
           "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _physical_index(self, pynvml):
    """NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
      ids = [x.strip() for x in vis.split(",") if x.strip()]
      if self.gpu < len(ids) and ids[self.gpu].isdigit():
        return int(ids[self.gpu])
    return self.gpu

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(",")])

  def stop(self):
    if getattr(self, "nvml_rows", None) is not None and self.proc is None and self.thread is not None:
      self.stop_flag = True
      self.thread.join(timeout=1)
      rows = self.nvml_rows
      reasons = sorted({r for row in rows for r in row[2]})
      return {"sm_mhz": statistics.median([r[0] for r in rows]) if rows else None,
              "sm_max_mhz": max([r[1] for r in rows]) if rows else None, "reasons": reasons,
              "samples": len(rows), "source": "nvml"}
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2]))
      except (ValueError, IndexError):
        continue
      for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                          "sw_power_cap"), r[5:9]):
        if v.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
            "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm
# ---------------------------------------------------------------------------
def cpu_threads():
  try:
    from threadpoolctl import threadpool_info
    n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
    return max(n) if n else (os.cpu_count() or 1)
  except Exception:  # pylint: disable=broad-except
    return os.cpu_count() or 1


def time_cpu_port(a, sample, reps=1, xs=None):
  """Times the oracle's matrix_inverse_pth_root (reference algorithm incl. its
  redundant mat_power products, DS:655-678) on `sample` matrices; roots/s.
  `xs`: the matrices to use (default: the numpy generator of the same statistics class)."""
  from oracle import numerics as N
  if xs is None:
    xs = make_statistics_numpy(sample, a.n, seed=1234)
  N.matrix_inverse_pth_root(xs[0][:64, :64].copy(), a.p)  # warm BLAS
  best = None
  iters, roots = [], []
  for _ in range(reps):
    iters, roots = [], []
    t0 = time.perf_counter()
    for b in range(sample):
      r, m = N.matrix_inverse_pth_root(xs[b], a.p, literal_mat_power=True)
      iters.append(m.inverse_pth_root_iters)
      roots.append((r, m))
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
  time_cpu_port.last_roots = roots
  return sample / best, best, iters


def parity_report(xs_dev, roots_dev, metrics_dev, oracle_roots, p):
  """Achieved parity of sampled bench matrices against the oracle (fp32 restatement of the
  reference): relative Frobenius distance of the roots, float64 residual max|X^p (A + eps I) - I|
  of both (torch float64 on the device as the checker), iteration counts."""
  import torch
  out = []
  for i, (want, wm) in enumerate(oracle_roots):
    a = xs_dev[i].double()
    n = a.shape[0]
    eye = torch.eye(n, dtype=torch.float64, device=a.device)
    d = a + 1e-6 * float(wm.max_eigen_value) * eye
    res = []
    for x in (roots_dev[i].double(), torch.as_tensor(want, device=a.device).double()):
      res.append(float((torch.linalg.matrix_power(x, p) @ d - eye).abs().max()))
    got = roots_dev[i].cpu().numpy()
    out.append({"rel_frobenius_vs_oracle": float(np.linalg.norm(got - want) / np.linalg.norm(want)),
                "residual_f64_ours": res[0], "residual_f64_oracle": res[1],
                "iters_ours": float(metrics_dev[i, 1]), "iters_oracle": float(wm.inverse_pth_root_iters),
                "error_ours": float(metrics_dev[i, 0]), "error_oracle": float(wm.inverse_pth_root_errors)})
  return {"samples": out,
          "max_rel_frobenius_vs_oracle": max(o["rel_frobenius_vs_oracle"] for o in out),
          "residual_no_worse_than_oracle": all(
              o["residual_f64_ours"] <= 2 * o["residual_f64_oracle"] + 1e-6 for o in out),
          "tolerance": "kappa ~1e6 class (SURVEY 8(c)): two fp32 solvers differ by ~ their float64 "
                       "residuals; tests assert rel <= 4 max(residuals) + 1e-3, iterations +-1"}


def run_reference(a):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  try:  # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host core
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=os.cpu_count() or 1)
  except Exception:  # pylint: disable=broad-except
    pass
  for _ in range(max(a.warmup, 0) and 1):
    time_cpu_port(a, 1)
  times = []
  for _ in range(a.steps):
    rps, dt, _ = time_cpu_port(a, a.cpu_sample)
    times.append(dt)
  dt = statistics.median(times)
  value = a.cpu_sample / dt
  cores = cpu_threads()
  sample = (f"{a.cpu_sample} statistics of the same workload per step, numpy/BLAS oracle port of "
            f"matrix_inverse_pth_root (JAX is not installable in this image)")
  line = {
      "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
      "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": workload_config(a, a.gpus),
      "run_info": {"note": "reference algorithm on host cores; bounded sample of "
                           f"{a.cpu_sample} statistics of the workload per step"},
      "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": sample},
      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
  """Pins this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host
  buffers of the end-to-end loop (first touch) live in the memory next to the GPU's PCIe root.
  Returns a short description for the JSON line; failures leave the affinity alone."""
  if os.environ.get("PC_BENCH_NUMA", "1") == "0":
    return "off"
  try:
    import pynvml
    pynvml.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
    bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    node = None
    for cand in (bus.lower(), bus.lower()[4:] if len(bus) > 12 else bus.lower()):
      path = f"/sys/bus/pci/devices/{cand}/numa_node"
      if os.path.exists(path):
        node = int(open(path).read().strip())
        break
    if node is None or node < 0:
      return f"unknown node for {bus}"
    cpus = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
      lo, _, hi = part.partition("-")
      cpus.update(range(int(lo), int(hi or lo) + 1))
    cpus &= os.sched_getaffinity(0)
    if not cpus:
      return f"node {node}: no allowed cpus"
    os.sched_setaffinity(0, cpus)
    return f"node {node} ({len(cpus)} cpus)"
  except Exception as e:  # pylint: disable=broad-except
    return f"unavailable ({type(e).__name__})"


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(a):
  import ctypes
  import torch
  import torch.distributed as dist
  from precondition_b200 import _lib, ops

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
  numa = bind_to_gpu_numa_node(local_rank)
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  lib = _lib.load()
  engine = {"auto": _lib.PC_ENGINE_AUTO, "simt": _lib.PC_ENGINE_SIMT_FP32,
            "tc6": _lib.PC_ENGINE_TC_BF16X6, "tc3": _lib.PC_ENGINE_TC_BF16X3,
            "fp16x3": _lib.PC_ENGINE_TC_FP16X3}[a.engine]

  n, B = a.n, a.batch
  xs = make_statistics_torch(B, n, seed=1000 + rank, device=dev)
  ps = torch.full((B,), a.p, dtype=torch.int32, device=dev)
  ps_host = np.full((B,), a.p, dtype=np.int32)
  roots = torch.empty_like(xs)
  metrics_buf = torch.empty((B, 5), dtype=torch.float32, device=dev)
  # Sub-batches: with several ranks the all-gather of the first half (DS:2876) runs on its own
  # stream underneath the solve of the second half; the solver call only enqueues.
  nsplit = a.split if a.split > 0 else 1
  nsplit = max(1, min(nsplit, B))
  bounds = [round(i * B / nsplit) for i in range(nsplit + 1)]
  parts = [(bounds[i], bounds[i + 1]) for i in range(nsplit) if bounds[i + 1] > bounds[i]]
  ws_parts = [torch.empty(ops.root_workspace_bytes(hi - lo, n, engine) + 256, dtype=torch.uint8,
                          device=dev) for lo, hi in parts]
  # all-gather of the roots (DS:2876): copy-engine pushes over NVLink peer memory
  # (precondition_b200/peer.py), NCCL as the fallback
  from precondition_b200 import peer
  gathered = ([peer.make_all_gather((hi - lo) * n * n * 4, None, dev) for lo, hi in parts]
              if world > 1 else None)
  comm_stream = torch.cuda.Stream(dev) if world > 1 else None
  gather_note = (f"all-gather ({gathered[0].kind}) of step k on a side stream under the solve of "
                 f"step k+1 (double-buffered roots; {len(parts)} sub-batch(es) per step)"
                 if world > 1 else "n/a (single gpu)")

  # The all-gather of step k runs on the copy engines underneath the solve of step k + 1 (no
  # SM is involved, precondition_b200/peer.py): two root buffers alternate, a buffer is reused
  # only after its gather has completed.  Every gather is inside the timed region (the region
  # ends with a synchronize, so the last one is fully exposed).
  gather_done = [None, None]
  step_no = [0]

  def solve(x_in, out):
    cur = torch.cuda.current_stream(dev)
    for k, (lo, hi) in enumerate(parts):
      ops.matrix_inverse_pth_root_batched(x_in[lo:hi], ps[lo:hi], None, engine=engine,
                                          out=out[lo:hi], metrics_out=metrics_buf[lo:hi],
                                          workspace=ws_parts[k], ps_host=ps_host[lo:hi])
      if world > 1:
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(comm_stream):
          comm_stream.wait_event(ev)
          gathered[k].all_gather(out[lo:hi])  # DS:2876
          gathered[k].release()               # (nothing reads the gathered copy in this bench)
    return out, metrics_buf

  roots_pp = [roots, torch.empty_like(roots)] if world > 1 else [roots, roots]

  def step():
    i = step_no[0] % 2
    step_no[0] += 1
    cur = torch.cuda.current_stream(dev)
    if world > 1 and gather_done[i] is not None:
      cur.wait_event(gather_done[i])  # the gather that read this buffer two steps ago
    m = solve(xs, roots_pp[i])[1]
    if world > 1:
      gather_done[i] = torch.cuda.Event()
      gather_done[i].record(comm_stream)
    return m

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(max(a.warmup, 3)):
    metrics = step()
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  stats = _lib.Stats()
  lib.pc_stats_reset(0)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for _ in range(a.steps):
    metrics = step()
  e1.record()
  barrier()
  lib.pc_stats_get(ctypes.byref(stats))
  launches = int(stats.kernel_launches) + (a.steps * len(parts) if world > 1 else 0)
  if lib.pc_root_mode():
    # graph mode: the library counts the graph's kernel nodes once per launch; the WHILE body
    # (re-init, G(p)-1 GEMM phases, control) repeats once per Newton iteration on the device
    body = 2 + (newton_gemms(a.p) - 1)
    launches += int(a.steps * len(parts) * body * max(float(metrics.cpu()[:, 1].max()) - 1, 0))
  ms = e0.elapsed_time(e1)
  t = torch.tensor([ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms = float(t.item())
  clocks = sampler.stop() if rank == 0 else None
  ms_per_step = ms / a.steps
  value = world * B / (ms_per_step * 1e-3)

  # ---- end-to-end through the C ABI with HOST buffers (pinned) ----
  host_in = torch.empty((B, n, n), dtype=torch.float32).pin_memory()
  host_in.copy_(xs.cpu())
  host_out = torch.empty((B, n, n), dtype=torch.float32).pin_memory()
  host_metrics = torch.empty((B, 5), dtype=torch.float32).pin_memory()
  dev_in = torch.empty_like(xs)

  # A double-buffered serving loop over the public API: the upload of step i+1 and the
  # download of step i-1 run on their own streams beside the solve of step i.  Every step's
  # H2D and D2H copies happen inside the timed region.
  cur = torch.cuda.current_stream(dev)
  s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
  dev_ins = [dev_in, torch.empty_like(xs)]
  dev_outs = [roots, torch.empty_like(xs)]
  ev_in = [torch.cuda.Event() for _ in range(2)]      # upload of slot landed
  ev_solved = [torch.cuda.Event() for _ in range(2)]  # solve that read/wrote slot finished
  ev_out = [torch.cuda.Event() for _ in range(2)]     # download of slot finished

  def upload(i):
    k = i % 2
    with torch.cuda.stream(s_h2d):
      s_h2d.wait_event(ev_solved[k])  # the previous user of this input slot is done
      dev_ins[k].copy_(host_in, non_blocking=True)
      ev_in[k].record(s_h2d)

  def e2e_loop(steps):
    s_h2d.wait_stream(cur)
    s_d2h.wait_stream(cur)
    upload(0)
    for i in range(steps):
      k = i % 2
      if i + 1 < steps:
        upload(i + 1)
      cur.wait_event(ev_in[k])
      cur.wait_event(ev_out[k])  # the previous download of this output slot is done
      r, m = solve(dev_ins[k], dev_outs[k])
      if world > 1:
        cur.wait_stream(comm_stream)  # the gathered copy of this step is complete
      ev_solved[k].record(cur)
      with torch.cuda.stream(s_d2h):
        s_d2h.wait_event(ev_solved[k])
        host_out.copy_(r, non_blocking=True)
        host_metrics.copy_(m, non_blocking=True)
        ev_out[k].record(s_d2h)
    cur.wait_stream(s_h2d)
    cur.wait_stream(s_d2h)

  e2e_loop(2)
  barrier()
  e0.record()
  e2e_loop(a.steps)
  e1.record()
  barrier()
  t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  e2e_ms = float(t.item()) / a.steps
  e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT,
         "h2d_bytes_per_step": int(B * n * n * 4), "d2h_bytes_per_step": int(B * n * n * 4 + B * 20),
         "ms_per_step": e2e_ms,
         "pipelining": "double-buffered: H2D of step i+1 and D2H of step i-1 overlap the solve of "
                       "step i (separate copy streams); all copies inside the timed region"}

  # ---- roofline of the dominant kernel (Newton-chain GEMM launches): one extra
  #      step with CUDA events around every GEMM launch on the launching stream ----
  lib.pc_stats_reset(1)
  metrics = step()  # untimed: the timing mode runs the host-polled driver, which warms up here
  torch.cuda.synchronize()
  lib.pc_stats_reset(1)
  for _ in range(3):  # totals over three calls: the launch average is taken over all of them
    metrics = step()
  torch.cuda.synchronize()
  lib.pc_stats_get(ctypes.byref(stats))
  lib.pc_stats_reset(0)
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except (OSError, ValueError):
    pass
  peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
  peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
  achieved = (stats.gemm_flops / (stats.gemm_ms * 1e-3) / 1e12) if stats.gemm_ms > 0 else 0.0
  traffic = None
  try:  # dram bytes per launch from the committed `ncu --set full` summary of this kernel
    tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    if tr.get("batch") == B and tr.get("n") == n:
      traffic = tr["dram_bytes_per_launch"]
  except (OSError, ValueError, KeyError):
    pass
  m_host = metrics.cpu().numpy()
  resolved = {1: "simt_fp32", 2: "tcgen05_bf16x6", 3: "tcgen05_bf16x3", 4: "tcgen05_fp16x3"}[
      lib.pc_resolve_engine(n, engine)]
  passes = {"simt_fp32": 1, "tcgen05_bf16x6": 6, "tcgen05_bf16x3": 3, "tcgen05_fp16x3": 3}[resolved]
  t_ = n // 128
  tile_frac = (t_ * (t_ + 1) / 2) / (t_ * t_) if (resolved != "simt_fp32" and t_ > 0) else 1.0
  roofline = {
      "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
      "frac": achieved / peak if peak else None, "traffic": traffic,
      "launches_timed": int(stats.gemm_launches),
      "algorithmic_flops_per_launch": stats.gemm_flops / max(int(stats.gemm_launches), 1),
      "kernel": "newton_chain_gemm", "engine": resolved,
      "timed_calls": 3,
      "algorithmic_flops_per_step": stats.gemm_flops / 3, "gemm_ms_per_step": stats.gemm_ms / 3,
      "gemm_share_of_step": stats.gemm_ms / 3 / ms_per_step if ms_per_step else None,
      "issued_passes": passes,
      # MMAs actually issued: `passes` products per k-step on the lower-triangular tiles only
      "issued_tile_fraction": tile_frac, "issued_tflops": achieved * passes * tile_frac,
      "peak_source": peak_src,
      "peak_burst": float(peaks.get("bf16_tflops", 0.0)) or None,
      "frac_of_burst": (achieved / float(peaks["bf16_tflops"])) if peaks.get("bf16_tflops") else None,
      "ceiling_frac": (1.0 / (passes * tile_frac)) if passes else None,
  }

  # ---- second half of the headline metric: full Shampoo step on BASELINE config 2
  #      (MLP 512->2048->512, block_size 128, SGD grafting, 1 GPU), through the
  #      optax-style API; steps >= 5 so the preconditioned path is active ----
  shampoo_step, sketchy, resnet_step, bert_step, sketchy_step = None, None, None, None, None
  small_block = None

  def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  if not a.no_step:
    resnet_step = time_resnet50_step(dev, world)  # every rank takes part (sharded roots)
    resnet_step["ms"] = max_over_ranks(resnet_step["ms"])
    if "sharded_vs_single_max_rel" in resnet_step:
      resnet_step["sharded_vs_single_max_rel"] = max_over_ranks(
          resnet_step["sharded_vs_single_max_rel"])
    if not a.no_big:
      bert_step = time_bert_large_step(dev, world)      # BASELINE config 4
      bert_step["ms"] = max_over_ranks(bert_step["ms"])
      sketchy_step = time_sketchy_step(dev, world)      # BASELINE config 5
      sketchy_step["ms"] = max_over_ranks(sketchy_step["ms"])
  tearfree_step = None
  if world == 1 and not a.no_step:
    shampoo_step = time_shampoo_step(dev)
    sketchy = time_sketchy_update(dev)
    try:
      tearfree_step = time_tearfree_step(dev)
    except Exception as e:  # pylint: disable=broad-except
      tearfree_step = {"error": f"{type(e).__name__}: {e}"[:200]}
    try:
      small_block = time_small_block_batch(dev)
    except RuntimeError as e:  # (not sm_100: the persistent solver needs tcgen05)
      small_block = {"error": str(e)[:200]}

  if rank == 0:
    cpu_baseline, parity = None, None
    if not a.no_cpu_baseline and world == 1:
      k = min(a.cpu_sample, B)
      rps, dt, iters = time_cpu_port(a, k, xs=xs[:k].cpu().numpy())
      cpu_baseline = {"value": rps, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                      "sample": f"the first {k} of this run's {B} statistics "
                                f"({dt:.2f} s; numpy/BLAS oracle port, JAX unavailable)",
                      "iters": iters}
      step()
      torch.cuda.synchronize()
      parity = parity_report(xs, roots, metrics_buf, time_cpu_port.last_roots, a.p)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, world),
        "run_info": {"engine": resolved, "host_numa_binding": numa,
                     "newton_iters_mean": float(m_host[:, 1].mean()),
                     "max_error": float(np.nanmax(m_host[:, 0])),
                     "root_mode": "cuda_graph_device_loop" if lib.pc_root_mode() else "host_polled",
                     "gather_overlap": gather_note},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches // 1,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
        "shampoo_step": shampoo_step, "sketchy_update": sketchy,
        "shampoo_step_resnet50": resnet_step, "shampoo_step_bert_large": bert_step,
        "sketchy_step": sketchy_step, "small_block_roots": small_block,
        "tearfree_step": tearfree_step,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def time_small_block_batch(dev, batch=64, n=128, p=4, reps=20):
  """BASELINE config 1: inverse 4th roots of a batch of 64 SPD 128 x 128 statistics (the
  reference's default block size) -- ms per batch on the persistent small-block solver and on
  the CUDA-core engine, iteration counts against the oracle for two of them."""
  import torch
  from oracle import numerics as N
  from oracle.gen_golden import ema_statistics, gen_symmetric_matrix
  from precondition_b200 import _lib, ops
  rng = np.random.default_rng(0)
  xs_h = np.stack([gen_symmetric_matrix(rng, n, 1e4) if i % 2 else ema_statistics(rng, n, 2 * n)
                   for i in range(batch)]).astype(np.float32)
  xs = torch.as_tensor(xs_h).to(dev)
  ps = torch.full((batch,), p, dtype=torch.int32, device=dev)
  ps_host = [p] * batch
  out = {"batch": batch, "n": n, "p": p, "unit": "ms/batch"}
  for name, engine in (("persistent_tcgen05", _lib.PC_ENGINE_TC_SMALL),
                       ("cuda_core_fp32", _lib.PC_ENGINE_SIMT_FP32)):
    roots, met = torch.empty_like(xs), torch.empty((batch, 5), device=dev)
    ws = torch.empty(ops.root_workspace_bytes(batch, n, engine) + 256, dtype=torch.uint8, device=dev)
    call = lambda: ops.matrix_inverse_pth_root_batched(xs, ps, None, engine=engine, out=roots,
                                                       metrics_out=met, workspace=ws,
                                                       ps_host=ps_host)
    for _ in range(3):
      call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name] = {"ms": ms, "roots_per_s": batch / ms * 1e3,
                 "newton_iters_mean": float(met[:, 1].mean()), "max_error": float(met[:, 0].max())}
    if name == "persistent_tcgen05":
      rel, its = [], []
      for b in (0, 1):
        want, wm = N.matrix_inverse_pth_root(xs_h[b], p)
        rel.append(float(np.linalg.norm(roots[b].cpu().numpy() - want) / np.linalg.norm(want)))
        its.append([float(met[b, 1]), float(wm.inverse_pth_root_iters)])
      out[name]["rel_frobenius_vs_oracle"] = rel
      out[name]["iters_ours_oracle"] = its
  return out


def time_shampoo_step(dev, steps=10, warm=6):
  """ms per `update` of distributed_shampoo on the synthetic 2-layer MLP of BASELINE
  config 2 (276 statistics of 128x128: 256 with p=4, 20 with p=2)."""
  import torch
  from precondition_b200 import distributed_shampoo as DS
  gen = torch.Generator(device=dev)
  gen.manual_seed(0)
  shapes = [(512, 2048), (2048,), (2048, 512), (512,)]
  params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
  opt = DS.distributed_shampoo(0.1, 128, graft_type=DS.GraftingType.SGD)
  state = opt.init(params)
  grads = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
           for _ in range(warm + steps)]
  for t in range(warm):  # same name as in the timed loop: the caching allocator reaches its steady
    upd, state = opt.update(grads[t], state, params)  # state (two update sets alive) before timing
  torch.cuda.synchronize()
  # one event pair per step: "ms" is the mean over the steps, the per-step list is kept so a
  # one-off stall (allocator, host scheduling) is visible instead of silently folded in
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
  evs[0].record()
  for t in range(warm, warm + steps):
    upd, state = opt.update(grads[t], state, params)
    evs[t - warm + 1].record()
  torch.cuda.synchronize()
  per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
  tm = torch.cat([st.training_metrics for st in state.stats if st.training_metrics is not None])
  return {"ms": sum(per_step) / steps, "ms_per_step_list": [round(x, 3) for x in per_step],
          "unit": "ms/step", "steps": steps,
          "config": "MLP 512->2048->512, block_size=128, SGD grafting, preconditioning_compute_steps=1",
          "statistics": int(tm.shape[0]), "newton_iters_mean": float(tm[:, 1].mean()),
          "max_root_error": float(tm[:, 0].max()),
          "update_finite": bool(all(torch.isfinite(u).all() for u in upd))}


def resnet50_shapes():
  """ResNet-50 v1.5 parameter shapes (HWIO conv kernels, BN scale / bias vectors, fc):
  BASELINE config 3."""
  shapes = [(7, 7, 3, 64), (64,), (64,)]
  cin = 64
  for width, blocks in ((64, 3), (128, 4), (256, 6), (512, 3)):
    for blk in range(blocks):
      cout = 4 * width
      for k, ci, co in ((1, cin, width), (3, width, width), (1, width, cout)):
        shapes += [(k, k, ci, co), (co,), (co,)]
      if blk == 0:  # projection shortcut
        shapes += [(1, 1, cin, cout), (cout,), (cout,)]
      cin = cout
  shapes += [(2048, 1000), (1000,)]
  return shapes


def _timed_updates(opt, state, params, grads, warm, steps):
  """Runs warm + steps updates; returns (per-step ms list, last updates, state)."""
  import torch
  upd = None
  for t in range(warm):  # same statement as the timed loop: the caching allocator reaches its
    upd, state = opt.update(grads[t], state, params)  # steady state before timing
  torch.cuda.synchronize()
  # one event pair per step: "ms" is the mean over the steps, the per-step list is kept so a
  # one-off stall (allocator, host scheduling) is visible instead of silently folded in
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
  evs[0].record()
  for t in range(warm, warm + steps):
    upd, state = opt.update(grads[t], state, params)
    evs[t - warm + 1].record()
  torch.cuda.synchronize()
  return [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)], upd, state


def _stat_sizes(state_stats):
  sizes = {}
  for st in state_stats:
    for x in st.statistics:
      k = int(x.shape[0])
      sizes[k] = sizes.get(k, 0) + 1
  return sizes


def time_resnet50_step(dev, world, steps=3, warm=3):
  """ms per `update` of distributed_shampoo on ResNet-50 shapes, block_size=1024,
  preconditioning_compute_steps=1 (BASELINE config 3); with more than one rank the
  preconditioner blocks are partitioned over the ranks and all-gathered (batch_axis_name),
  and the same steps are repeated UNSHARDED on this rank to report
  `sharded_vs_single_max_rel` (max over parameters of max|u_sharded - u_single| / max|u_single|
  after the last step, and the same for every preconditioner)."""
  import torch
  from precondition_b200 import distributed_shampoo as DS
  gen = torch.Generator(device=dev)
  gen.manual_seed(3)  # same parameters / gradients on every rank (replicated state)
  shapes = resnet50_shapes()
  params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
  opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1,
                               batch_axis_name="batch" if world > 1 else None)
  state = opt.init(params)
  grads = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
           for _ in range(warm + steps)]
  per_step, upd, state = _timed_updates(opt, state, params, grads, warm, steps)
  tm = torch.cat([st.training_metrics for st in state.stats if st.training_metrics is not None])
  sizes = _stat_sizes(state.stats)
  out = {"ms": sum(per_step) / steps, "ms_per_step_list": [round(x, 3) for x in per_step],
         "unit": "ms/step", "steps": steps, "n_gpus": world,
         "config": "ResNet-50 shapes, block_size=1024, preconditioning_compute_steps=1, "
                   "blocks sharded over the ranks" if world > 1 else
                   "ResNet-50 shapes, block_size=1024, preconditioning_compute_steps=1",
         "parameters": int(sum(p.numel() for p in params)), "statistics": int(tm.shape[0]),
         "statistics_of_1024": sizes.get(1024, 0),
         "max_root_error": float(tm[:, 0].max()),
         "update_finite": bool(all(torch.isfinite(u).all() for u in upd))}
  if world > 1:
    single = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1)
    sstate = single.init(params)
    _, supd, sstate = _timed_updates(single, sstate, params, grads, warm, steps)
    rel = 0.0
    for u, v in zip(upd, supd):
      rel = max(rel, float((u - v).abs().max() / v.abs().max().clamp_min(1e-30)))
    prel = 0.0
    for a_, b_ in zip(state.stats, sstate.stats):
      for x, y in zip(a_.preconditioners, b_.preconditioners):
        prel = max(prel, float((x - y).abs().max() / y.abs().max().clamp_min(1e-30)))
    out["sharded_vs_single_max_rel"] = max(rel, prel)
    out["sharded_vs_single_updates_max_rel"] = rel
    out["sharded_vs_single_preconditioners_max_rel"] = prel
    del single, sstate, supd
    try:  # the pjit layout (DS:2162-2583): statistics stored / updated / solved by their owner only
      zero = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1,
                                    shard_optimizer_states=True, num_devices_for_pjit=world)
      zstate = zero.init(params).init_fn(params)  # DS:2585-2625: init returns an InitFnState
      zper, zupd, zstate = _timed_updates(zero, zstate, params, grads, warm, steps)
      zrel = 0.0
      for u, v in zip(upd, zupd):
        zrel = max(zrel, float((u - v).abs().max() / u.abs().max().clamp_min(1e-30)))
      out["shard_optimizer_states"] = {"ms": sum(zper) / steps,
                                       "ms_per_step_list": [round(x, 3) for x in zper],
                                       "updates_vs_replicated_max_rel": zrel}
    except Exception as e:  # pylint: disable=broad-except
      out["shard_optimizer_states"] = {"error": f"{type(e).__name__}: {e}"[:200]}
  return out


def bert_large_shapes():
  """BERT-large (340M) parameter shapes (BASELINE config 4): 24 layers of Q/K/V [1024,16,64],
  attention output [16,64,1024], FFN 1024x4096 / 4096x1024, biases and LayerNorm vectors,
  embeddings (the 30522 x 1024 table is skipped by DS:2627-2629), pooler."""
  shapes = [(30522, 1024), (512, 1024), (2, 1024), (1024,), (1024,)]
  for _ in range(24):
    shapes += [(1024, 16, 64), (16, 64)] * 3
    shapes += [(16, 64, 1024), (1024,), (1024,), (1024,)]
    shapes += [(1024, 4096), (4096,), (4096, 1024), (1024,), (1024,), (1024,)]
  shapes += [(1024, 1024), (1024,)]
  return shapes


def time_bert_large_step(dev, world, steps=2, warm=2):
  """BASELINE config 4: BERT-large shapes, block_size=2048, best_effort_memory_usage_reduction
  (int16 statistics / preconditioners with extracted diagonal, int8 momenta),
  preconditioning_compute_steps=1; blocks sharded over the ranks when world > 1.  Also reports
  the relative Frobenius distance of one sampled 1024^2 preconditioner from the oracle's root of
  the same (dequantised) statistic -- the int16 storage quantum is part of it."""
  import torch
  from precondition_b200 import distributed_shampoo as DS
  gen = torch.Generator(device=dev)
  gen.manual_seed(4)
  shapes = bert_large_shapes()
  params = [torch.randn(s, generator=gen, device=dev) * 0.02 for s in shapes]
  opt = DS.distributed_shampoo(0.1, 2048, preconditioning_compute_steps=1,
                               best_effort_memory_usage_reduction=True, batch_axis_name="batch")
  state = opt.init(params)
  # two gradient sets, reused alternately (the full list would hold 1.3 GB per step)
  gsets = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes] for _ in range(2)]
  grads = [gsets[t % 2] for t in range(warm + steps)]
  per_step, upd, state = _timed_updates(opt, state, params, grads, warm, steps)
  leaves = state.stats
  tm = torch.cat([st.training_metrics for st in leaves if st.training_metrics is not None])
  sizes = _stat_sizes(leaves)
  out = {"ms": sum(per_step) / steps, "ms_per_step_list": [round(x, 3) for x in per_step],
         "unit": "ms/step", "steps": steps, "n_gpus": world,
         "config": "BERT-large shapes, block_size=2048, best_effort_memory_usage_reduction "
                   "(int16 statistics + preconditioners, int8 momenta), "
                   "preconditioning_compute_steps=1",
         "parameters": int(sum(p.numel() for p in params)), "statistics": int(tm.shape[0]),
         "statistics_by_size": {str(k): v for k, v in sorted(sizes.items())},
         "max_root_error": float(tm[:, 0].max()), "newton_iters_mean": float(tm[:, 1].mean()),
         "update_finite": bool(all(torch.isfinite(u).all() for u in upd))}
  if int(os.environ.get("RANK", "0")) == 0:
    from oracle import numerics as N
    st = next(s_ for s_ in leaves if s_.statistics and s_.statistics[0].shape[0] == 1024)
    a_ = st.statistics[0].to_float().cpu().numpy()
    got = st.preconditioners[0].to_float().cpu().numpy()
    p_ = DS.Preconditioner(shapes[leaves.index(st)], 2048, 4096,
                           True).exponent_for_preconditioner()
    want, wm = N.matrix_inverse_pth_root(a_, p_)
    out["sampled_root_rel_frobenius_vs_oracle"] = float(
        np.linalg.norm(got - want) / np.linalg.norm(want))
    out["sampled_root_iters_ours_oracle"] = [float(st.training_metrics[0, 1]),
                                             float(wm.inverse_pth_root_iters)]
  return out


def time_tearfree_step(dev, steps=3, warm=3):
  """ms per `update` of the tearfree front-end (precondition_b200.tearfree.optimizer.tearfree) on
  a transformer-block-shaped parameter set (hidden 1024, FFN 4096), two configurations of
  TF/second_order.py: blocked Shampoo at block_size 256 (eigh-based pseudo-inverse roots every
  step) and Sketchy at rank 128; RMSProp grafting, Nesterov momentum (the defaults)."""
  import torch
  from precondition_b200.tearfree import grafting, momentum, optimizer, second_order, shampoo
  from precondition_b200.tearfree import sketchy
  gen = torch.Generator(device=dev)
  gen.manual_seed(6)
  shapes = [(1024, 1024)] * 4 + [(1024, 4096), (4096, 1024), (1024,), (4096,), (1024,)]
  params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
  grads = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
           for _ in range(warm + steps)]
  out = {"unit": "ms/step", "steps": steps, "parameters": int(sum(p.numel() for p in params)),
         "config": "4 x 1024^2 + 1024x4096 + 4096x1024 + 3 vectors; RMSProp grafting, Nesterov 0.9"}
  for name, so in (
      ("shampoo_block256", second_order.Options(
          merge_dims=1024, shampoo_options=shampoo.Options(block_size=256))),
      ("sketchy_rank128", second_order.Options(
          merge_dims=1024, second_order_type=second_order.SecondOrderType.SKETCHY,
          sketchy_options=sketchy.Options(rank=128)))):
    tx = optimizer.tearfree(0.1, optimizer.TearfreeOptions(second_order_options=so))
    state = tx.init(params)
    per_step, upd, state = _timed_updates(tx, state, params, grads, warm, steps)
    out[name] = {"ms": sum(per_step) / steps, "ms_per_step_list": [round(x, 3) for x in per_step],
                 "update_finite": bool(all(torch.isfinite(u).all() for u in upd))}
  return out


def time_sketchy_step(dev, world, steps=2, warm=2):
  """BASELINE config 5: Sketchy / frequent-directions branch on 8 parameters of 4096 x 4096,
  block_size=4096, rank 256 (16 statistics), sketch updates sharded over the ranks."""
  import torch
  from precondition_b200 import distributed_shampoo as DS
  gen = torch.Generator(device=dev)
  gen.manual_seed(5)
  shapes = [(4096, 4096)] * 8
  params = [torch.randn(s, generator=gen, device=dev) * 0.02 for s in shapes]
  opt = DS.distributed_shampoo(0.1, 4096, compression_rank=256, frequent_directions=True,
                               reuse_preconditioner=True, statistics_compute_steps=1,
                               preconditioning_compute_steps=1,
                               batch_axis_name="batch" if world > 1 else None)
  state = opt.init(params)
  gsets = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes] for _ in range(2)]
  grads = [gsets[t % 2] for t in range(warm + steps)]
  per_step, upd, state = _timed_updates(opt, state, params, grads, warm, steps)
  pk = state.stats[0].preconditioners[0]
  v = pk[:, :256]
  return {"ms": sum(per_step) / steps, "ms_per_step_list": [round(x, 3) for x in per_step],
          "unit": "ms/step", "steps": steps, "n_gpus": world,
          "config": "8 x (4096 x 4096), block_size=4096, compression_rank=256, "
                    "frequent_directions, reuse_preconditioner, statistics_compute_steps="
                    "preconditioning_compute_steps=1",
          "statistics": 16, "statistics_per_gpu": (16 + world - 1) // world,
          "eigvec_orthogonality_error": float(
              (v.T @ v - torch.eye(256, device=dev)).abs().max()),
          "update_finite": bool(all(torch.isfinite(u).all() for u in upd))}


def time_sketchy_update(dev, d=4096, rank=256, batch=2, steps=2):
  """ms per batched Sketchy / frequent-directions sketch update at BASELINE config 5's
  per-GPU share (16 statistics of 4096 x 4096, rank 256, over 8 GPUs = 2 per GPU)."""
  import torch
  from precondition_b200 import ops
  gen = torch.Generator(device=dev)
  gen.manual_seed(5)
  u = torch.linalg.qr(torch.randn(d, d, generator=gen, device=dev))[0]
  spec = torch.cat([torch.logspace(0, -1.5, rank + 64, device=dev),
                    torch.full((d - rank - 64,), 0.01, device=dev)])
  xs = torch.stack([(u * spec) @ torch.randn(d, d, generator=gen, device=dev) / d**0.5
                    for _ in range(batch)]).contiguous()
  sketch = torch.zeros((batch, d, rank + 2), device=dev)
  ps = [4] * batch
  for _ in range(2):  # warm the sketch up
    sketch, _ = ops.fd_update_root_batched(xs, sketch, ps, rank, decay=0.999)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(steps):
    sketch, _ = ops.fd_update_root_batched(xs, sketch, ps, rank, decay=0.999)
  e1.record()
  torch.cuda.synchronize()
  v = sketch[0, :, :rank]
  orth = float((v.T @ v - torch.eye(rank, device=dev)).abs().max())
  return {"ms": e0.elapsed_time(e1) / steps, "unit": "ms/update", "d": d, "rank": rank,
          "batch": batch, "config": "frequent_directions sketch update, 4096x4096 blocks, rank 256",
          "eigvec_orthogonality_error": orth, "has_zeros": float(sketch[0, -1, -2])}


def main():
  a = parse_args()
  if a.impl == "reference":
    run_reference(a)
  else:
    run_ours(a)


if __name__ == "__main__":
  main()
