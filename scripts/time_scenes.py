"""bench.scenes_block alone (4096 GC-shaped scenes, 10 frames by default): for a launch list / timing of the scene-batched
rollout.  python scripts/time_scenes.py [scenes] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
print(bench.scenes_block(torch, None, torch.device("cuda", 0), 1, 0, S_total=S, steps=steps))
