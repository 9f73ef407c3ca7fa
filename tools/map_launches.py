"""One mapping iteration (996 rays, stage color, dense voxel grads + colour-decoder grads) a few times: run under
ncu --metrics gpu__time_duration.sum to list its launches (tools/gpu_round4.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import scene_util as su
from gpu_util import make_renderer
from nice_slam_b200.steps import IterationContext
import bench
dev = torch.device("cuda")
sc = su.load_scenes()["room0"]
renderer, c, dec = make_renderer(sc, su.make_grids(sc, "soft"), su.load_decoders("soft"), dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 996
ro, rd, dirs, gd, gc = [t.to(dev) for t in bench.make_batch(sc, n, 101)]
ctx = IterationContext(renderer, n, "color", dev, kind="map", grad_grids=("grid_middle", "grid_fine", "grid_color"), grad_decoders=("color",))
for _ in range(4):
    ctx.run(c, dec, ro, rd, gd, gc.float())
torch.cuda.synchronize()
print("loss", float(ctx.loss))
