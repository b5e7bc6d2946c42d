import sys, torch, numpy as np, os
sys.path.insert(0, os.getcwd())
from multi_spatialmllm_b200 import synth, _lib
from multi_spatialmllm_b200.engine import DeviceScene
s = synth.make_scene(seed=0, n_points=200_000, n_frames=512, n_objects=12)
for ro in (None, "morton"):
    sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, reorder=ro)
    sc.render_depth(1)
    st = torch.zeros(2, dtype=torch.int64, device='cuda')
    _lib.lib.msp_debug_set_cull_stats(st.data_ptr())
    sc.project_visibility(); torch.cuda.synchronize()
    _lib.lib.msp_debug_set_cull_stats(None)
    tot = sc.n * sc.F
    print(ro, 'point-frames', tot, 'candidate groups x32 lanes', int(st[0])*32, f'({int(st[0])*32/tot:.3f})', 'exact lanes', int(st[1]), f'({int(st[1])/tot:.5f})', 'visible', int(sc.mask_bool().sum())/tot, flush=True)
