"""Turns gpurun_out/{launches.csv, prof_gemm.ncu-rep, prof_gn.ncu-rep, profile_ops.csv, bench.json} into the small
tracked summaries under profiles/ (ncu itself runs here without a GPU: `ncu -i ... --page raw --csv`)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_src_fp16_dst_fp32.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]


def ncu_raw(rep):
  txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(txt)))
  hdr, units = rows[0], rows[1]
  out = []
  for r in rows[2:]:
    d = {}
    for i, h in enumerate(hdr):
      if h in KEEP:
        d[h] = f"{r[i]} {units[i]}".strip()
    out.append(d)
  return out


def main():
  os.makedirs(OUT, exist_ok=True)
  lines = []
  p = os.path.join(SRC, "launches.csv")
  if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 5 and r[0].strip('"').isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
      name = r[4].split("(")[0]
      try:
        ns = float(r[-1].replace(",", ""))
      except ValueError:
        continue
      unit = r[-2]
      us = ns / 1000.0 if unit == "ns" else (ns if unit in ("us", "usecond") else ns * 1000.0)
      agg[name][0] += 1
      agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    lines.append(f"## ncu launch list of ONE network evaluation (deep NCSN++, batch 256): {sum(v[0] for v in agg.values())} launches, "
                 f"{tot/1000:.2f} ms serialized (cold cache; compare shares, not absolutes)\n")
    lines.append("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      lines.append(f"| {k} | {v[0]} | {v[1]:.0f} | {100*v[1]/tot:.1f}% |")
    lines.append("")
  p = os.path.join(SRC, "gemm_traffic.csv")
  if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 5 and r[0].strip('"').isdigit()]
    per = collections.defaultdict(dict)
    for r in rows:
      val = float(r[-1].replace(",", ""))
      unit = r[-2]
      if r[-3] == "gpu__time_duration.sum":
        val = val / 1000.0 if unit == "ns" else (val if unit in ("us", "usecond") else val * 1000.0)
      else:
        val = val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
      per[r[0]][r[-3]] = val
      per[r[0]]["name"] = r[4].split("(")[0]
    n = len(per)
    rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in per.values())
    wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in per.values())
    us = sum(v.get("gpu__time_duration.sum", 0.0) for v in per.values())
    json.dump({"kernel": "conv_gemm_umma_kernel (all launches of one network evaluation, deep NCSN++, batch 256)",
               "launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd + wr) / max(n, 1),
               "sum_duration_us": us, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (tools/profile.sh)"},
              open(os.path.join(OUT, f"{TAG}_gemm_traffic.json"), "w"), indent=1)
    lines.append(f"## DRAM traffic of the {n} conv_gemm_umma launches of one evaluation: read {rd/1e9:.2f} GB, write {wr/1e9:.2f} GB, "
                 f"{(rd+wr)/max(n,1)/1e6:.1f} MB per launch; algorithmic minimum = operands + outputs (see DESIGN.md)\n")
    import shutil
    shutil.copy(p, os.path.join(OUT, f"{TAG}_gemm_traffic.csv"))
  for rep, title in (("prof_gemm256.ncu-rep", "conv_gemm_umma_kernel<256,0,1,*> (six consecutive N=256 launches of the 16x16 level: 3x3 convs as CTA pairs and the K=256 attention GEMMs; ncu --set full)"),
                     ("prof_gemm128.ncu-rep", "conv_gemm_umma_kernel<128,0,2,2,true> (3x3 conv 128->128 @32x32: 256-row halo tiles, CTA pairs; ncu --set full)"),
                     ("prof_gn.ncu-rep", "gn_apply_kernel (ncu --set full)"),
                     ("prof_attn.ncu-rep", "attn256_kernel (fused QK^T -> softmax -> P.V of one 16x16 attention block; ncu --set full)")):
    p = os.path.join(SRC, rep)
    if os.path.exists(p):
      lines.append(f"## {title}\n")
      for d in ncu_raw(p):
        lines.append("- " + "; ".join(f"{k}={v}" for k, v in d.items()))
      lines.append("")
  p = os.path.join(SRC, "profile_ops.csv")
  if os.path.exists(p):
    rows = list(csv.DictReader(open(p)))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
      k = (r["kind"], r["H"], r["N"], r["K"], r["block_n"])
      agg[k][0] += 1; agg[k][1] += float(r["ms_per_forward"]); agg[k][2] += float(r["gflop"])
    tot = sum(v[1] for v in agg.values())
    lines.append(f"## per-op CUDA-event timing of one evaluation inside bench.py (eager launches): {tot:.2f} ms\n")
    lines.append("kind: 0 stem, 1 groupnorm(stats+apply), 2 conv_gemm (tcgen05), 3 head, 4 im2col, 5 transpose_v, 6 small attention, 8 fused attention (tcgen05)\n")
    lines.append("| kind | H=W | N (C_out) | K | block_n | launches | ms | share | TFLOP/s |\n|---|---|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      tf = f"{v[2]/v[1]:.0f}" if v[2] > 0 and v[1] > 0 else ""
      lines.append(f"| {k[0]} | {k[1]} | {k[2]} | {k[3]} | {k[4]} | {v[0]} | {v[1]:.3f} | {100*v[1]/tot:.1f}% | {tf} |")
    lines.append("")
    import shutil
    shutil.copy(p, os.path.join(OUT, f"{TAG}_per_op.csv"))
  p = os.path.join(SRC, "bench.json")
  if os.path.exists(p):
    line = open(p).read().strip().splitlines()[-1]
    try:
      json.loads(line)
      open(os.path.join(OUT, f"{TAG}_bench.json"), "w").write(line + "\n")
      lines.insert(0, f"## bench.py line of the same build\n\n```\n{line}\n```\n")
    except Exception:
      pass
  p = os.path.join(SRC, "launches.csv")
  if os.path.exists(p):
    import shutil
    shutil.copy(p, os.path.join(OUT, f"{TAG}_ncu_launches.csv"))
  open(os.path.join(OUT, f"{TAG}_summary.md"), "w").write(f"# Profile summary {TAG}\n\n" + "\n".join(lines) + "\n")
  print("wrote", os.path.join(OUT, f"{TAG}_summary.md"))


if __name__ == "__main__":
  main()
