"""Quick check of the persistent panel chain against the oracle (panel level), one subprocess per mode so that a hang
in one mode is reported instead of stalling the run:  python tools/chain_check.py [timeout_s]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, os, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, oracle
from gpu_util import panel_factor
for (m, n, lam, pw) in [(300, 64, 0, 64), (2048, 128, 0, 128), (5000, 256, 128, 128), (20000, 128, 0, 128), (32768, 128, 0, 128)]:
    A = oracle.uniform_matrix(m, n, 7 * m + n)
    P0 = oracle.pack(A)
    if lam:
        P0 = oracle.householder_panel(P0.copy(), 0, lam)
    Pref = oracle.householder_panel(P0.copy(), lam, pw)
    Wref, Yref = oracle.wy_factors(Pref, lam, pw)
    t0 = time.time()
    P, Y, W, T = panel_factor(P0.copy(), lam, pw)
    dt = time.time() - t0
    sc = np.abs(Pref).max()
    print(f"  {m}x{n} lam={lam} pw={pw}: |P-Pref|/max={np.abs(P - Pref).max() / sc:.2e} |Y-Yref|={np.abs(Y - Yref).max():.2e} "
          f"|W-Wref|={np.abs(W - Wref).max():.2e}  ({dt:.2f} s)", flush=True)
''' % (ROOT, ROOT)

tmo = int(sys.argv[1]) if len(sys.argv) > 1 else 90
for name, env in [("classic (MPQR_NO_CHAIN=1)", {"MPQR_NO_CHAIN": "1"}), ("chain + gate kernels", {"MPQR_GATE_KERNEL": "1"}),
                  ("chain, default ordering (stream wait + kernel-side post)", {})]:
    e = dict(os.environ); e.update(env)
    print(name, flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], env=e, timeout=tmo, capture_output=True, text=True)
        print(r.stdout, end="")
        if r.returncode != 0:
            print("  FAILED rc=%d\n%s" % (r.returncode, r.stderr[-1500:]))
    except subprocess.TimeoutExpired as ex:
        print((ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else (ex.stdout or ""), end="")
        print(f"  TIMEOUT after {tmo} s (hang)", flush=True)
