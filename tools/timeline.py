"""Kernel timeline of ONE factorisation (CUPTI through torch.profiler: start / duration / stream of every kernel,
concurrent streams included).  python tools/timeline.py m,n,r,prec [t0_ms t1_ms]
Prints per-kernel-name totals and, for the window [t0, t1) ms after the first kernel, every kernel in start order
(columns: start us, duration us, stream, grid, name).  The full list goes to gpurun_out/timeline_<m>x<n>.csv."""
import sys, os, json, collections, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import mixedprecisionblockqr_b200 as pkg

spec = sys.argv[1].split(",")
tsqr = spec[0] == "tsqr"   # tsqr,m,n : mpqr_tsqr_device (R + thin Q)
if tsqr:
    spec = [spec[1], spec[2], "128", "fp32"]
m, n, r, prec = int(spec[0]), int(spec[1]), int(spec[2]), spec[3]
w0 = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
w1 = float(sys.argv[3]) if len(sys.argv) > 3 else 21.5
lda = (n + 7) // 8 * 8
A = torch.zeros(m + 1, lda, device="cuda")
st = torch.cuda.current_stream().cuda_stream
if tsqr:
    class _T:   # same interface as a plan
        def __init__(self):
            self.Q = torch.zeros(m, n, device="cuda"); self.R = torch.zeros(n, n, device="cuda")
        def factor(self, a, ld, s):
            pkg.tsqr(a, ld, m, n, self.Q.data_ptr(), n, self.R.data_ptr(), n, s)
        def close(self):
            pass
    plan = _T()
else:
    plan = pkg.BlockQR(m, n, r, precision=prec)
for it in range(2):
    pkg.fill_uniform(A.data_ptr(), lda, n, 0, m, 0, n, 1234, st)
    plan.factor(A.data_ptr(), lda, st)
    torch.cuda.synchronize()
pkg.fill_uniform(A.data_ptr(), lda, n, 0, m, 0, n, 1234, st)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.factor(A.data_ptr(), lda, st)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "mpqr_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ev)
print(f"{m}x{n} r={r} {prec}: {len(ev)} kernels, span {(end - t0) / 1e3:.2f} ms (under the profiler)")


def short(name):
    name = name.replace("mpqr::", "").replace("(anonymous namespace)::", "")
    return name.split("(")[0][:70]


tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = short(e["name"])
    tot[k][0] += 1
    tot[k][1] += e["dur"]
print("totals by kernel (count, total ms, avg us):")
for k, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {c:6d} {d / 1e3:9.2f} {d / c:8.1f}  {k}")
# activity per millisecond: kernels started, summed kernel time, streams in use
nbin = int((end - t0) / 1e3) + 1
if nbin <= 64:
    print("per ms: kernels started / summed kernel ms / streams / top kernel")
    for b in range(nbin):
        sel = [e for e in ev if b <= (e["ts"] - t0) / 1e3 < b + 1]
        if not sel:
            print(f"  {b:3d}: -")
            continue
        byname = collections.Counter()
        for e in sel:
            byname[short(e["name"])] += e["dur"]
        print(f"  {b:3d}: {len(sel):5d} {sum(e['dur'] for e in sel) / 1e3:7.2f} {len({e.get('args', {}).get('stream') for e in sel}):3d}  {byname.most_common(1)[0][0]}")
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/timeline_{m}x{n}.csv", "w") as f:
    f.write("start_us,dur_us,stream,grid,name\n")
    for e in ev:
        a = e.get("args", {})
        f.write(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{a.get('stream')},{'x'.join(map(str, a.get('grid', [])))},{short(e['name'])}\n")
print(f"window [{w0}, {w1}) ms:")
for e in ev:
    s = (e["ts"] - t0) / 1e3
    if w0 <= s < w1:
        a = e.get("args", {})
        print(f"  {e['ts'] - t0:10.1f} {e['dur']:8.1f}  s{a.get('stream'):<4} g{'x'.join(map(str, a.get('grid', []))):<10} {short(e['name'])}")
plan.close()
