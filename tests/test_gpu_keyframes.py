"""GPU: the device-resident keyframe store against the REAL Mapper.keyframe_selection_overlap (tests/golden/keyframe_overlap.pt,
generated from the unmodified reference by tests/make_golden.py) and against the oracle (oracle/keyframes.py)."""
import os

import numpy as np
import pytest
import torch

import scene_util as su

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case():
    return torch.load(os.path.join(su.GOLDEN, "keyframe_overlap.pt"), map_location="cpu", weights_only=False)


def test_overlap_counts_and_selection_match_the_real_mapper():
    from nice_slam_b200.keyframes import KeyframeStore
    from oracle import keyframes as okf
    case = _case()
    sc = su.load_scenes()[case["scene"]]
    cam = sc["cam"]
    depth, color = su.make_frame(sc, case["frame_seed"])
    store = KeyframeStore(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], DEV, capacity=4)      # grows twice
    for k, c2w in enumerate(case["keyframe_c2w"]):
        store.append(5 * k, color, depth, c2w)
    ro, rd, gd = case["rays_o"], case["rays_d"], case["gt_depth"]
    n = ro.shape[0] * 16
    counts = store.overlap_counts(ro, rd, gd).cpu()
    want = torch.tensor([round(p * n) for p in case["percent_inside"]])
    # the projection is float32 / float64 arithmetic in another summation order than numpy's BLAS: a point within rounding of an image edge may
    # fall on the other side -- none does on this fixture, and at most 2 of 1600 are tolerated
    assert int((counts - want).abs().max()) <= 2, (counts, want)
    assert torch.equal(counts == 0, want == 0)                  # keyframes without overlap are exactly those of the reference
    # oracle on the same inputs
    pts = okf.overlap_points(ro, rd, gd)
    mine = [okf.percent_inside(pts, c2w, cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])[1] for c2w in case["keyframe_c2w"]]
    assert int((counts - torch.tensor(mine)).abs().max()) <= 2
    if torch.equal(counts, want):                                # identical percent_inside -> identical selection under the same numpy seed
        rng = np.random.RandomState(case["numpy_seed"])
        sel = store.select_overlap(ro, rd, gd, case["k"], rng=rng)
        assert [int(x) for x in sel] == case["selected"]


def test_window_samples_come_from_the_resident_images():
    from nice_slam_b200.keyframes import KeyframeStore
    sc = su.load_scenes()["room0"]
    cam = sc["cam"]
    store = KeyframeStore(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], DEV, capacity=2)
    frames = [su.make_frame(sc, 20 + k) for k in range(5)]
    for k, (d, c) in enumerate(frames):
        store.append(k, c, d, su.make_pose(sc, 20 + k))
    g = torch.Generator().manual_seed(3)
    slots = [4, 0, 2]
    pi = torch.randint(cam["W"], (3, 166), generator=g); pj = torch.randint(cam["H"], (3, 166), generator=g)
    od, oc = store.sample(slots, pi, pj)
    for f, s in enumerate(slots):
        d, c = frames[s]
        assert torch.equal(od[f].cpu(), d[pj[f], pi[f]].float())
        assert torch.equal(oc[f].cpu(), c[pj[f], pi[f]].float())
    # pose write-back after BA changes the overlap of that keyframe only
    ro, rd, _, gd, _ = __import__("bench").make_batch(sc, 100, 5)
    before = store.overlap_counts(ro, rd, gd).cpu()
    flip = su.make_pose(sc, 22).clone(); flip[:3, :3] = flip[:3, :3] @ torch.diag(torch.tensor([-1.0, 1.0, -1.0]))
    store.update_pose(2, flip)
    after = store.overlap_counts(ro, rd, gd).cpu()
    assert torch.equal(before[[0, 1, 3, 4]], after[[0, 1, 3, 4]]) and int(after[2]) == 0 and int(before[2]) > 0
