"""Where the host time of tables.visibility_arrow_table goes (one 200 k-point x 375-frame scene, cProfile, top entries)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multi_spatialmllm_b200 import engine, synth, tables  # noqa: E402

s = synth.make_scene(seed=5, n_points=200_000, n_frames=375, n_objects=12)
sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align)
sc.render_depth(1)
sc.project_visibility()
tables.visibility_arrow_table(sc, s.scene_id, s.image_ids)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    tables.visibility_arrow_table(sc, s.scene_id, s.image_ids)
print("s per table:", (time.perf_counter() - t0) / 3)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    tables.visibility_arrow_table(sc, s.scene_id, s.image_ids)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
