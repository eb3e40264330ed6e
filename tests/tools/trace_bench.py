"""Dev tool (GPU box): time the nearest-hit query alone (csrc/rt_trace_bench.cuh) on the recorded rays of a workload.
usage: trace_bench.py [C3] [max_rays]   (env RT_B200_WQ_BURST / _T_LEAF / _T_PEND / _T_FIN tune the state machine)"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ray-tracer-s8_b200"))
import rt_b200 as rt
from rt_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 64_000_000
cfg = scenes.CONFIGS[name]
sp, tr = scenes.config_scene(name)
ctx = rt.Context(0)
sc = ctx.scene(sp, tr)
p = rt.make_params(cfg["width"], cfg["height"], spp=cfg["spp"], max_bounces=cfg["max_bounces"], intersector=2)
for with_big, srt in ((True, False), (False, False), (False, True)):
    r = ctx.trace_bench(sc, p, cap, with_big, srt)
    n = r["rays"]
    print(f"{name} with_big={int(with_big)} sorted={int(srt)} rays={n} mismatches={r['mismatches']} "
          f"while-while {r['ms_while_while']:.3f} ms ({n / r['ms_while_while'] / 1e3:.0f} Mrays/s)  "
          f"state machine {r['ms_state_machine']:.3f} ms ({n / r['ms_state_machine'] / 1e3:.0f} Mrays/s)")
