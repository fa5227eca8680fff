"""gpurun_out/r02_* (tools/profile_r02.sh) -> tracked summaries under profiles/: the launch list, the GEMM DRAM-traffic
table (+ JSON read by bench.py), and one line per `ncu --set full` capture with the roofline-relevant counters."""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = PEAKS.get("hbm_gbs", 6551.4)
PFX = os.environ.get("PFX", "r02")     # PFX=r02f: the captures of tools/profile_r02_final.sh -> profiles/r02f_ncu_summary.md
WANT = collections.OrderedDict([
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__cluster_size", "cluster"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")])


def to_bytes(v, unit):
  v = float(v.replace(",", ""))
  return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
  v = float(v.replace(",", ""))
  return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)


def main():
  commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
  lines = [f"# {PFX} ncu captures (commit {commit}; `tools/profile_r02.sh` / `tools/profile_r02_final.sh`, deep NCSN++ batch 256, `--set full --clock-control none`)", "",
           f"HBM peak for the fractions: {HBM} GB/s (MEASURED_PEAKS.json).  DRAM bytes are per launch; cold-cache, serialised launches.", "",
           "| capture | kernel | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | frac of HBM peak | tensor pipe % | XU pipe % | issue % | L2 hit % | regs | grid x cluster |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
  for f in sorted(os.listdir(SRC)):
    if not (f.startswith(PFX + "_") and f.endswith("_raw.csv")):
      continue
    rows = list(csv.reader(open(os.path.join(SRC, f))))
    if len(rows) < 3:
      continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
      d = {}
      for i, h in enumerate(hdr):
        if h in WANT:
          d[WANT[h]] = (r[i], units[i])
        if h == "Kernel Name":
          d["name"] = r[i]
      t = to_us(*d["time"])
      rd, wr = to_bytes(*d["dram_rd"]), to_bytes(*d["dram_wr"])
      gbs = (rd + wr) / (t * 1e-6) * 1e-9
      g = lambda k: d.get(k, ("", ""))[0]
      name = d.get("name", "?").split("(")[0][-70:]
      lines.append(f"| {f[len(PFX) + 1:-8]} | `{name}` | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {gbs / HBM:.2f} | {g('tensor_pipe_pct')} | "
                   f"{g('xu_pipe_pct')} | {g('issue_pct')} | {g('l2_hit_pct')} | {g('regs')} | {g('grid')} x {g('cluster')} |")
      os.makedirs(OUT, exist_ok=True)
  # launch list: per-kernel totals of one evaluation
  p = os.path.join(SRC, PFX + "_launches.csv")
  if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 5 and r[0].strip('"').isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
      name = r[4].split("(")[0]
      try:
        us = to_us(r[-1], r[-2])
      except ValueError:
        continue
      agg[name][0] += 1; agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    what = "launch list of one evaluation" if PFX == "r02" else \
        "the first launches of `python bench.py --steps 2 --warmup 1` (150 time-embedding `dense_kernel` launches of sampler creation, then the first evaluations with the update inside the head convolution: no `cld_step` launches)"
    lines += ["", f"## {what} ({len(rows)} launches, {tot / 1e3:.2f} ms serialised; shares are what to compare with bench.py)", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      lines.append(f"| `{k[-90:]}` | {v[0]} | {v[1]:.0f} | {v[1] / tot:.3f} |")
    with open(os.path.join(OUT, PFX + "_ncu_launches.csv"), "w") as o:
      o.write(open(p).read())
  p = os.path.join(SRC, PFX + "_gemm_traffic.csv")
  if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    hdr = next(r for r in rows if "Metric Name" in r)
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    per = collections.defaultdict(dict)
    for r in rows:
      if len(r) > iv and r[0].strip('"').isdigit():
        per[r[iid]][r[im]] = (r[iv], r[iu])
    n, tot_b, tot_t = 0, 0.0, 0.0
    for _, m in per.items():
      if "dram__bytes_read.sum" in m:
        tot_b += to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
        tot_t += to_us(*m["gpu__time_duration.sum"])
        n += 1
    if n:
      json.dump({"commit": commit, "launches": n, "dram_bytes_per_launch": tot_b / n, "dram_bytes_per_evaluation": tot_b,
                 "avg_launch_us_under_ncu": tot_t / n, "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every conv_gemm_umma launch of one evaluation (tools/profile_r02.sh)"},
                open(os.path.join(OUT, PFX + "_gemm_traffic.json"), "w"), indent=1)
      lines += ["", f"## GEMM family DRAM traffic: {n} launches, {tot_b / 1e9:.2f} GB per evaluation, {tot_b / n / 1e6:.1f} MB per launch (profiles/{PFX}_gemm_traffic.json)"]
      with open(os.path.join(OUT, PFX + "_gemm_traffic.csv"), "w") as o:
        o.write(open(p).read())
  open(os.path.join(OUT, PFX + "_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
  print("\n".join(lines[:40]))


if __name__ == "__main__":
  main()
