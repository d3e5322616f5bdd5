"""GPU debug aid: three frames through the tracker-only pipeline (pyramid by TMA unless XIVO_PYRDOWN_TMA=0); run under compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xivo_b200 import pyxivo, sim, synth

cfg = sim.load_cfg(os.path.join(ROOT, "xivo_b200", "cfg", "vio_640x480.json"))
rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 240, int(sys.argv[2]) if len(sys.argv) > 2 else 320
cfg["camera_cfg"].update(rows=rows, cols=cols, fx=cols * 0.43, fy=cols * 0.43, cx=cols / 2, cy=rows / 2)
cfg["tracker_cfg"].update(num_features_min=60, num_features_max=80)
cfg["message_buffer_size"] = 0
canvas = synth.texture_canvas(rows, cols, seed=5, pad=64)
frames = [synth.frame_from_canvas(canvas, rows, cols, (3 * k, 2 * k), noise_seed=50 + k, pad=32) for k in range(3)]
b = pyxivo.Batch(cfg, n_seq=2, max_groups=4, max_features=14, tracker_only=True)
for k, img in enumerate(frames):
    b.visual_meas(k * 40_000_000, [img, img], tracker_only=True)
    ids, xy, st = b.tracked_features(0)
    print("frame", k, "tracks", len(ids), flush=True)
b.close()
print("ok")
