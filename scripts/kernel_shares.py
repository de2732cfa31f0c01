"""Warm (no ncu, no cache flush) event-bracketed shares of the config 2 step: Schur updates, panel chain, assembly.
    python scripts/kernel_shares.py [side]          (GPU box)
Each kind is measured in its own process-wide pass (NGT_B200_PROFILE_KIND is read once), so the script re-executes itself."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if os.environ.get("NGT_SHARES_CHILD"):
    from kmc_rates_b200 import synth
    from kmc_rates_b200._capi import Handle
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 316
    n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B)
    for _ in range(3):
        h.timed_steps(1)
    best = None
    for _ in range(5):
        ms, fl, nl = h.profile_schur()
        pass_ms = h.stats()["ms_profiled_pass"]
        if best is None or ms < best[0]:
            best = (ms, nl, pass_ms)
    print("%-8s %7.3f ms in %4d bracketed launches, profiled pass %.3f ms (graph step %.3f ms)" %
          (os.environ.get("NGT_B200_PROFILE_KIND", "update"), best[0], best[1], best[2], h.timed_steps(10) / 10))
else:
    for kind in ("update", "panel", "assemble"):
        env = dict(os.environ, NGT_SHARES_CHILD="1", NGT_B200_PROFILE_KIND=kind)
        subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, check=True)
