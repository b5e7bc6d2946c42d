"""configs[4] (stress) step on one scene, for `ncu --metrics gpu__time_duration.sum` launch lists and CUDA-event timing:
500k uniform random points, 128 frames; K2 render + u16 conversion + K1 + pair stage.  REORDER=morton sorts the points."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multi_spatialmllm_b200 import engine, synth  # noqa: E402

reorder = os.environ.get("REORDER") or None
F = int(os.environ.get("SP_FRAMES", "128"))
s = synth.make_stress_scene(seed=7000, n_points=500_000, n_frames=F)
sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, device="cuda:0", reorder=reorder)
pairs = torch.arange(F, dtype=torch.int32, device="cuda").view(-1, 2)
sel = torch.randint(-2**31, 2**31 - 1, (F // 2,), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(4):
    ev[0].record(); sc.render_depth(1, chunk=F)
    ev[1].record(); sc.project_visibility()
    ev[2].record(); sc.pair_annotate(pairs, sel, sel_mode=1)
    ev[3].record()
torch.cuda.synchronize()
print(json.dumps({"reorder": reorder, "frames": F, "k2_render_ms": ev[0].elapsed_time(ev[1]), "k1_ms": ev[1].elapsed_time(ev[2]),
                  "pair_ms": ev[2].elapsed_time(ev[3]), "visible_bits": int(sc.mask_bool().sum())}))
