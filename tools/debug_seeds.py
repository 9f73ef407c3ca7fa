"""Debug helper: stand-alone tracking seeds kernel vs torch for a few n / flags."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from nice_slam_b200 import _lib
from oracle import torch_port as tp
L = _lib.lib(); DEV = "cuda"
for n in (200, 513, 777, 5000):
    for hd in (0, 1):
        g = torch.Generator().manual_seed(9)
        depth = torch.rand(n, generator=g, dtype=torch.float64) * 3
        var = torch.rand(n, generator=g, dtype=torch.float64) * 0.1
        rgb = torch.rand(n, 3, generator=g)
        gt = torch.rand(n, generator=g) * 3
        gt[::13] = 0
        gt_rgb = torch.rand(n, 3, generator=g, dtype=torch.float64)
        res = (gt.double() - depth).abs() / (var + 1e-10).sqrt()
        med = res.median()
        m = (gt > 0) & ((res < 10 * med) if hd else torch.ones_like(gt, dtype=torch.bool))
        want = res[m].sum() + 0.5 * (gt_rgb - rgb.double()).abs()[m].sum()
        gD = torch.full((n,), -7.0, dtype=torch.float64, device=DEV); gC = torch.empty(n, 3, device=DEV); lo = torch.full((1,), -1.0, dtype=torch.float64, device=DEV)
        ws = torch.zeros(L.nsb_tracking_seeds_workspace(n), dtype=torch.uint8, device=DEV)
        t = [x.to(DEV) for x in (depth, var, rgb, gt, gt_rgb)]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.nsb_tracking_seeds(*[C.c_void_p(x.data_ptr()) for x in t], n, 0.5, hd, 1, None, 0, C.c_void_p(gD.data_ptr()), C.c_void_p(gC.data_ptr()),
                                  C.c_void_p(lo.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), st)
        torch.cuda.synchronize()
        r2 = ws.view(torch.float64)[:n].cpu()
        print("n %5d hd %d rc %d loss %.6f want %.6f  res ok %s  nonzero gD %d (want %d) median %.6f" %
              (n, hd, rc, float(lo), float(want), bool(torch.allclose(r2, res)), int((gD != 0).sum()), int(m.sum()), float(med)))
