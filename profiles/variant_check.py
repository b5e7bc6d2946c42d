"""One-process check of a library variant (MSP_LIB): K1 with fused counts on a config-2-shaped scene (200 k points, 512 frames),
time per launch by CUDA events and a hash of the masks + counts — equal hashes across variants = bit-identical output."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multi_spatialmllm_b200 import engine, synth  # noqa: E402

NF = int(os.environ.get('VC_FRAMES', '512'))
s = synth.make_scene(seed=5, n_points=200_000, n_frames=NF, n_objects=12)
sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=s.num_objects,
                        image_hw=s.image_hw, depth_hw=s.depth_hw)
sc.render_depth(1)
mask, counts = sc.project_visibility_objects()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    mask, counts = sc.project_visibility_objects()
e1.record()
torch.cuda.synchronize()
h = hashlib.sha256(mask.cpu().numpy().tobytes() + counts.cpu().numpy().tobytes()).hexdigest()[:16]
print(json.dumps({"lib": os.path.basename(os.environ.get("MSP_LIB", "default")), "frames": NF, "k1_ms": round(e0.elapsed_time(e1) / 10, 4),
                  "visible_bits": int(sum(bin(int(x) & 0xffffffff).count("1") for x in mask[:4].cpu().numpy().ravel())), "sha": h}), flush=True)
