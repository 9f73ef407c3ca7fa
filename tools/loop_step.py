"""Mapping iteration vs the native mapper loop step (996 rays, stage color; bench.py extra_workloads' first two entries), quick form."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import scene_util as su
from gpu_util import make_renderer
from nice_slam_b200.steps import IterationContext
from nice_slam_b200.mapping import FusedMappingLoop
import bench
dev = torch.device("cuda")
sc = su.load_scenes()["room0"]
renderer, c, dec = make_renderer(sc, su.make_grids(sc, "soft"), su.load_decoders("soft"), dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_steps(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


n = 996
ro, rd, dirs, gd, gc = [t.to(dev) for t in bench.make_batch(sc, n, 101)]
gcf = gc.float()
ctx = IterationContext(renderer, n, "color", dev, kind="map", grad_grids=("grid_middle", "grid_fine", "grid_color"), grad_decoders=("color",))
print("mapping iteration      %.4f ms" % time_steps(lambda: ctx.run(c, dec, ro, rd, gd, gcf), 100))
depth1, _ = su.make_frame(sc, 1)
loop = FusedMappingLoop(renderer, {k: v.clone() for k, v in c.items()}, copy.deepcopy(dec), su.make_pose(sc, 1), depth1.to(dev))
ctxm = loop._context(n, "color", dev)
print("masked iteration       %.4f ms" % time_steps(lambda: ctxm.run(loop.c, loop.dec, ro, rd, gd, gcf, w_color=loop.w_color), 100))
lr = dict(decoders=0.005, middle=0.005, fine=0.005, color=0.005)
print("native loop step       %.4f ms" % time_steps(lambda: loop.iteration("color", ro, rd, gd, gcf, lr), 100))
