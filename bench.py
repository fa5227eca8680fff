#!/usr/bin/env python
"""Benchmark of the gDDIM sampling hot path (BASELINE.json metric: CIFAR-10 32x32 images/sec at 50 NFE,
deis_order=2, CLD, deep NCSN++ of cld_jax/configs/accr_dcifar10_config.py, batch 256 per GPU).

  python bench.py --gpus N --steps K --warmup W            ours (under torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    CPU restatement of the reference sampler (oracle/)

A "step" is one full sampler call on one batch: 49 multistep network evaluations + 1 denoising evaluation.
`value` times device-resident inputs; `e2e` goes through the reference-shaped public API
(sampling.get_sampling_fn -> psampler) with host buffers, copies inside the timed region.
The reference's own JAX/XLA path cannot run in this image (jax/flax absent): the reference arm
and `cpu_baseline` time the torch-CPU fp32 restatement of the same sampler ("port"), never labelled JAX.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = "CIFAR10 32x32 images/sec @ 50 NFE deis_order=2 (CLD, deep NCSN++)"
GFLOP_PER_IMG_EVAL = {"deep": 37.168, "ddpmpp": 21.707}
TRAFFIC_FILE = "r02f_gemm_traffic.json"                         # ncu DRAM pass of the GEMM family (tools/profile.sh)


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--net", default="deep", choices=["deep", "ddpmpp"])
  ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
  ap.add_argument("--nfe", type=int, default=50)
  ap.add_argument("--order", type=int, default=2)
  ap.add_argument("--cpu-batch", type=int, default=0, help="images per CPU baseline sample (0 = auto)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--workload", default="cld", choices=["cld", "blur"],
                  help="cld = BASELINE configs 2/4/5 (default: config 2); blur = config 3 (order-0 DDIM in DCT space)")
  ap.add_argument("--image-size", type=int, default=32, help="256 = BASELINE config 5 geometry (use --batch 8)")
  ap.add_argument("--profile-csv", default="", help="write the per-op timing table of one profiled step here")
  ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                  help="BASELINE.json configs[k-1] preset (per-GPU sizes; 2 = the default the driver runs): "
                       "1 = deep CLD, batch 4, NFE 10, order 0; 2 = deep CLD, batch 256, NFE 50, order 2; "
                       "3 = blur, batch 256, NFE 50; 4 = deep CLD, order 3, 256 per GPU (global 2048 on 8 GPUs); "
                       "5 = deep CLD at 256x256, order 2, 8 per GPU (global 64 on 8 GPUs)")
  args = ap.parse_args()
  preset = {1: dict(batch=4, nfe=10, order=0), 2: dict(batch=256, nfe=50, order=2),
            3: dict(workload="blur", batch=256, nfe=50), 4: dict(batch=256, nfe=50, order=3),
            5: dict(image_size=256, batch=8, nfe=50, order=2)}.get(args.config, {})
  for k, v in preset.items():
    setattr(args, k, v)
  return args


def make_cfg(name, workload="cld", image_size=32):
  from gddim_b200 import configs
  if workload == "blur":
    return configs.blur_ddpm_deep_cifar10(1.0)
  cfg = configs.cld_accr_dcifar10() if name == "deep" else configs.cld_ddpmpp_cifar10()
  cfg.data.image_size = image_size
  return cfg


class ClockSampler:
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.idx, self.rows, self.proc = gpu_index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      pass
    sm, mx, reasons = [], [], set()
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2]))
        for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
          if r[col].lower().startswith("active"):
            reasons.add(name)
      except Exception:
        continue
    busy = sorted(sm)[len(sm) // 2:] if sm else []
    return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_sample(cfg, params, batch, nfe, order, threads):
  """Times the oracle (CPU restatement of the reference sampler) on `batch` images; returns images/s."""
  import torch
  from oracle import cld as oc
  from oracle import ncsnpp as on
  torch.set_num_threads(threads)
  sde = oc.from_config(cfg)
  eps_fn = oc.make_eps_fn(sde, on.make_net_fn(params, cfg))
  u = oc.prior_sampling(np.random.default_rng(0), (batch, 32, 32, 3)).astype(np.float32)
  t0 = time.perf_counter()
  oc.deis_sampler(sde, eps_fn, u, nfe, order, denoising=True, dtype=np.float32)
  dt = time.perf_counter() - t0
  return batch / dt, dt


def run_reference(args):
  """--impl reference: the reference sampler's CPU restatement on all host threads; rank 0 only."""
  rank = int(os.environ.get("RANK", 0))
  if rank != 0:
    return
  # parameter inventory from the ORACLE's own walk (oracle.ncsnpp.collect_specs) and the plain-python generator: this arm
  # never loads libgddim_b200.so (gddim_b200.configs / params are data-only modules)
  from gddim_b200 import params as gparams
  from oracle import ncsnpp as on
  cfg = make_cfg(args.net)
  params = gparams.generate(on.collect_specs(cfg, cld=True), seed=1234, nondegenerate=True)
  threads = os.cpu_count() or 1
  batch = args.cpu_batch or 8
  for _ in range(max(args.warmup, 0)):
    cpu_port_sample(cfg, params, batch, min(args.nfe, 4), args.order, threads)     # short warm-up (thread pools, caches)
  t = 0.0
  for _ in range(args.steps):
    _, dt = cpu_port_sample(cfg, params, batch, args.nfe, args.order, threads)
    t += dt
  v = batch * args.steps / t
  line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": f"CLD CIFAR10 32x32, NFE={args.nfe}, deis_order={args.order}, net={args.net} "
                                 f"(bounded sample: {batch} image(s) per step on the host CPU)"},
          "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                           "sample": f"{batch} image(s) x {args.nfe} NFE per step, {args.steps} steps, torch-CPU fp32 "
                                     "restatement of cld_jax sampler (JAX not installable offline)"},
          "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  emit(line)


def run_ours(args):
  import torch
  from gddim_b200 import dist as gdist
  from gddim_b200 import net
  from gddim_b200.cld import sampling, sde_lib

  rank, local_rank, world = gdist.init_process_group()
  assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
  torch.cuda.set_device(local_rank)
  B, nfe, order = args.batch, args.nfe, args.order
  blur, S = args.workload == "blur", args.image_size
  default_workload = (not blur) and S == 32
  cfg = make_cfg(args.net, args.workload, S)
  if blur:
    from gddim_b200.blur import sampling, sde_lib
    cfg.sampling.method, cfg.sampling.nfe = "order0", nfe
  else:
    cfg.sampling.method, cfg.sampling.nfe, cfg.sampling.deis_order = "deis", nfe, order
  model = net.ScoreNet(cfg, cld=not blur)
  # parameters: generated on rank 0, one NCCL broadcast (run_lib.py:711 replicate)
  if rank == 0:
    params = model.init_params(seed=1234, nondegenerate=True)
  else:
    params = {n: s[0] for n, s in model.specs().items()}
  params = gdist.broadcast_params(params, src=0)
  if rank != 0:
    model.set_params(params)
  sde = sde_lib.from_config(cfg)
  inv = lambda x: (x + 1.) / 2.
  psampler = sampling.get_sampling_fn(cfg, sde, model, None, inv)
  core = psampler.core

  # one global prior draw, sliced per rank (sampling.py:235)
  rng = np.random.default_rng(0)
  shape = (world * B, S, S, 3)
  if blur:
    u_glob = rng.standard_normal(shape).astype(np.float32)      # y ~ N(0, 1) in DCT space (blur sde_lib.py:128-130)
  else:
    u_glob = np.stack([rng.standard_normal(shape), rng.standard_normal(shape) / np.sqrt(cfg.model.m_inv)],
                      axis=-1).astype(np.float32)        # x ~ N(0,1), v ~ N(0, 1/m_inv)  (sde_lib.py:270-274)
  u_host = np.ascontiguousarray(gdist.shard(u_glob, rank, world))
  u_pin = torch.from_numpy(u_host).pin_memory()
  u_dev = u_pin.cuda()
  stream = torch.cuda.current_stream()

  def step_dev():
    return core.run(model, B, u_dev)

  for _ in range(max(args.warmup, 3)):
    x = step_dev()[0]
  torch.cuda.synchronize()

  # ---- timed region: device-resident inputs -------------------------------------------------------------
  clocks = ClockSampler(local_rank)
  gdist.barrier()
  torch.cuda.synchronize()
  l0 = core.launch_count()
  if rank == 0:
    clocks.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(stream)
  for _ in range(args.steps):
    x = step_dev()[0]
  e1.record(stream)
  torch.cuda.synchronize()
  gdist.barrier()
  ms = gdist.max_over_ranks(e0.elapsed_time(e1))
  clk = clocks.stop() if rank == 0 else None
  launches = core.launch_count() - l0
  value = world * B * args.steps / (ms * 1e-3)
  assert torch.isfinite(x).all()

  # ---- e2e: public API, host buffers (pinned), copies inside the timed region --------------------------------
  u_np = u_pin.numpy()[None]                       # (n_dev=1, B, 32, 32, 3, 2) view of pinned memory
  outs = psampler(None, model, B, u=u_np)          # warm-up of the host path
  gdist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  e0.record(stream)
  for _ in range(args.steps):
    outs = psampler(None, model, B, u=u_np)
  e1.record(stream)
  torch.cuda.synchronize()
  wall = time.perf_counter() - t0
  gdist.barrier()
  ms_e2e = gdist.max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
  e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
  h2d = u_host.nbytes
  d2h = sum(o.nbytes for o in outs[:-1])           # (xs, vs) for CLD, (xs,) for blur; the last item is nfe

  # ---- roofline of the dominant kernel (conv_gemm_umma) from one profiled step ------------------------------
  roof, roof_hbm = None, None
  if rank == 0:
    model.set_profile(True)
    core.run(model, B, u_dev)
    torch.cuda.synchronize()
    ms_kind, gemm_flops, gemm_launches = model.get_profile()
    norm_bytes = model.get_profile_norm_bytes()
    if args.profile_csv:
      os.makedirs(os.path.dirname(os.path.abspath(args.profile_csv)), exist_ok=True)
      model.dump_profile(args.profile_csv)
    model.set_profile(False)
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
      pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = gemm_flops / (ms_kind["conv_gemm"] * 1e-3) * 1e-12 if ms_kind["conv_gemm"] > 0 else 0.0
    tot = sum(ms_kind.values())
    # DRAM bytes per launch of the same kernel family from the committed ncu pass (profiles/, tools/profile.sh)
    traffic = None
    try:
      tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
      traffic, traffic_commit = tj["dram_bytes_per_launch"], tj.get("commit", "unrecorded")
    except Exception:
      traffic_commit = None
    # ---- HBM-bound kernel families against the measured copy bandwidth ---------------------------------------------
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    gn_ms = ms_kind["groupnorm"]
    upd_ms, upd_bytes = C.c_double(), C.c_double()
    from gddim_b200 import _lib as glib
    glib.check(glib.lib().gddim_sampler_time_update(core.handle(model, B), B, 200, C.byref(upd_ms), C.byref(upd_bytes),
                                                    stream.cuda_stream), "gddim_sampler_time_update")
    torch.cuda.synchronize()
    roof_hbm = {
        "peak": hbm_peak, "unit": "GB/s",
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if peaks else "fallback 6.5 TB/s",
        "groupnorm": {"kernels": "gn_coef / gn_apply<*> / gn_small / gn_stats (all OP_NORM launches of one step)",
                      "bytes": norm_bytes, "ms": gn_ms, "achieved": norm_bytes / (gn_ms * 1e-3) * 1e-9 if gn_ms > 0 else 0.0,
                      "frac": (norm_bytes / (gn_ms * 1e-3) * 1e-9) / hbm_peak if gn_ms > 0 else 0.0,
                      "bytes_def": "fp32 source read once (+ once more where the statistics are not produced by the GEMM epilogue), fp16 outputs written once"},
        "update": {"kernel": "blur_step_kernel" if blur else "cld_step_c3_kernel",
                   "bytes_per_launch": upd_bytes.value, "avg_launch_ms": upd_ms.value,
                   "achieved": upd_bytes.value / (upd_ms.value * 1e-3) * 1e-9 if upd_ms.value > 0 else 0.0,
                   "frac": (upd_bytes.value / (upd_ms.value * 1e-3) * 1e-9) / hbm_peak if upd_ms.value > 0 else 0.0,
                   "in_timed_step": bool(blur),
                   "note": ("one launch per step behind the network evaluation" if blur else
                            "the timed step applies this update inside the head convolution's epilogue (epi_head_update, same "
                            "per-pixel arithmetic, no separate launch); the standalone kernel measured here serves stochastic "
                            "and traced calls"),
                   "bytes_def": "(order+3) x 24576 B per image (CLD) / 4 x 12288 B per image (blur), SURVEY.md 8d; 200 launches back to back in one CUDA-event pair on rotating buffer sets (>= 400 MB in total, 3x the L2), so every launch reads from HBM"}}
    roof = {"bound": "tensor", "kernel": "conv_gemm_umma_kernel (all conv3x3 / 1x1 / NIN GEMM launches; the fused QK^T-softmax-PV kernel is the separate 'attention' family)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
            "flop_per_launch": gemm_flops / max(gemm_launches, 1),
            "traffic": traffic, "traffic_source": f"profiles/{TRAFFIC_FILE} (ncu dram bytes, measured at commit {traffic_commit})",
            "launches": gemm_launches, "avg_launch_ms": ms_kind["conv_gemm"] / max(gemm_launches, 1),
            "share_of_step": ms_kind["conv_gemm"] / tot if tot else None,
            # these kernels also execute the GroupNorm + swish of most layers (fused epilogue), so `frac` is not comparable
            # with a GEMM-only number; FLOPs over (GEMM + remaining GroupNorm passes) time is comparable across builds
            "frac_conv_plus_groupnorm": (gemm_flops / ((ms_kind["conv_gemm"] + ms_kind["groupnorm"]) * 1e-3) * 1e-12) / peak
            if ms_kind["conv_gemm"] > 0 else None,
            "ms_by_kernel_family": {k: round(v_, 3) for k, v_ in ms_kind.items()}}

  gdist.barrier()
  if rank != 0:
    return
  # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ------------------------------
  cpu = None
  if world == 1 and not args.no_cpu_baseline and default_workload:
    threads = os.cpu_count() or 1
    cb = args.cpu_batch or 8
    v_cpu, dt = cpu_port_sample(cfg, model.params, cb, nfe, order, threads)
    cpu = {"value": v_cpu, "unit": "images/s", "cores": threads, "kind": "port",
           "sample": f"{cb} image(s) x {nfe} NFE, same net/sampler, torch-CPU fp32 restatement of the reference "
                     f"sampler ({dt:.1f} s); JAX/XLA itself is not installable offline"}
  if default_workload or (blur and args.net == "deep"):
    flop_img = (37.152 if blur else GFLOP_PER_IMG_EVAL[args.net]) * 1e9 * nfe
  else:
    # other geometries: 2*M*N*K of every GEMM launch of the profiled step (includes the 6-channel stem/head padding)
    flop_img = (roof["flop_per_launch"] * roof["launches"] / B) if roof else 0.0
  if blur:
    wl = (f"Blur-diffusion CIFAR10 32x32, sigma_blur_max=1.0, batch={B}/GPU, NFE={nfe}, order0 (DDIM in DCT space), "
          "net=ddpm_deep_cifar10 (deep NCSN++, 107.6M params)")
    metric = f"CIFAR10 32x32 images/sec @ {nfe} NFE order0 (blur diffusion, deep NCSN++)"
  else:
    wl = (f"CLD {'CIFAR10 ' if S == 32 else ''}{S}x{S}, batch={B}/GPU, NFE={nfe}, deis_order={order}, "
          f"net={'accr_dcifar10 (deep NCSN++, 107.6M params)' if args.net == 'deep' else 'ddpmpp_cifar10'}")
    metric = METRIC if (S == 32 and order == 2 and nfe == 50 and args.net == "deep") else \
        f"{S}x{S} images/sec @ {nfe} NFE deis_order={order} (CLD, {args.net})"
  line = {"metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f16 operands, f32 accumulate/trunk", "data": "synthetic",
          "config": {"workload": wl + ", random-init (non-degenerate) weights, Gaussian prior",
                     "global_batch": world * B, "parallelism": f"dp{world} (batch-sharded, no per-step collective)",
                     "l2": f"working set >> L2: {model.workspace_bytes() / 2**30:.2f} GiB of activations+weights per evaluation"},
          "tensor_frac_end_to_end": value * flop_img / (world * 1e12 * (roof["peak"] if roof else 1400.0)),
          "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
          "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_hbm": roof_hbm}
  if cpu is not None:
    line["cpu_baseline"] = cpu
  emit(line)


_JSON_FD = None


def emit(line):
  """The one JSON line of the contract, on the process's ORIGINAL stdout (see main)."""
  data = (json.dumps(line) + "\n").encode()
  if _JSON_FD is None:
    sys.stdout.write(data.decode()); sys.stdout.flush()
  else:
    os.write(_JSON_FD, data)


def main():
  # NCCL prints its version banner (and NCCL_DEBUG output) on stdout -- from C, behind Python's back; the contract is ONE
  # JSON line there.  Everything this process (and the libraries it loads) writes to fd 1 goes to stderr instead; only
  # emit() writes to the original stdout.
  global _JSON_FD
  os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
  sys.stdout.flush()
  _JSON_FD = os.dup(1)
  os.dup2(2, 1)
  args = parse()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)
    try:
      import torch.distributed as dist
      if dist.is_initialized():
        dist.destroy_process_group()
    except Exception:
      pass


if __name__ == "__main__":
  main()
