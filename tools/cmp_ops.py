"""Compare two per-op CSVs (tools/op_times.py) by (op suffix, H, N, K): count, total ms.  usage: cmp_ops.py a.csv b.csv"""
import collections, csv, sys
def load(p):
  d = collections.OrderedDict()
  for r in csv.DictReader(open(p)):
    suf = r["op"].split("/")[-1]
    k = (suf, r["H"], r["N"], r["K"])
    e = d.setdefault(k, [0, 0.0])
    e[0] += 1; e[1] += float(r["ms_per_forward"])
  return d
a, b = load(sys.argv[1]), load(sys.argv[2])
keys = list(a.keys()) + [k for k in b if k not in a]
ta = tb = 0.0
for k in keys:
  ea, eb = a.get(k, [0, 0.0]), b.get(k, [0, 0.0])
  ta += ea[1]; tb += eb[1]
  if max(ea[1], eb[1]) > 0.02:
    print(f"{k[0]:22s} H={k[1]:>3s} N={k[2]:>4s} K={k[3]:>5s}  A: n={ea[0]:3d} {ea[1]:7.3f} ms ({1e3*ea[1]/max(ea[0],1):6.1f} us)   B: n={eb[0]:3d} {eb[1]:7.3f} ms ({1e3*eb[1]/max(eb[0],1):6.1f} us)")
print(f"total A {ta:.3f} ms  B {tb:.3f} ms")
