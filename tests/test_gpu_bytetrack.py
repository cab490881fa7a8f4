"""GPU parity of the upstream tracker plug-in (ByteTrack): IoU / fused-score cost matrices bit-exact, linear assignment equal to
the restated lapjv optimum (and to the reference-generated cases), BYTETracker.update over the scripted scene against the
reference's own output (ids / scores / classes / detection indices bit-exact, boxes within the fp32 Kalman tolerance)."""
import os

import numpy as np
import pytest

import b200dt  # noqa: F401
from oracle import byte_tracker as obt

from golden_common import bytetrack_script

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
BYTE_PARAMS = {"default": {}, "nofuse": dict(track_high_thresh=0.4, track_low_thresh=0.15, new_track_thresh=0.5, track_buffer=8,
                                             match_thresh=0.7, fuse_score=False)}


def _boxes(g, n, span=300.0):
    c = g.uniform(0, span, (n, 2))
    s = g.uniform(4, 60, (n, 2))
    return np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)


def test_iou_cost_bit_exact_vs_reference_arithmetic():
    """b2_iou_cost == 1 - bbox_ioa(a, b, iou=True) (and fuse_score on top) in the reference's float32 operation order, bit for
    bit, including disjoint, identical and degenerate (zero-area) boxes; batched problems with ragged sizes leave the rest alone."""
    import torch

    from b200dt import _lib, byte_tracker as bt

    g = np.random.default_rng(0)
    a, b = _boxes(g, 37), _boxes(g, 53)
    a[3] = b[5]                                        # identical pair
    a[4] = [10, 10, 10, 30]                            # zero-width box
    b[6] = [500, 500, 520, 520]                        # disjoint from everything
    sc = g.uniform(0.05, 1, 53).astype(np.float32)
    ref = 1 - obt.bbox_iou_f32(a, b)
    got = bt.iou_distance(list(a), list(b))
    assert got.dtype == np.float32 and np.array_equal(got, ref)
    fused = 1 - (1 - ref) * sc[None].repeat(37, 0)
    assert np.array_equal(bt.iou_distance(list(a), list(b), sc), fused)
    assert bt.iou_distance([], list(b)).shape == (0, 53) and bt.iou_distance(list(a), []).shape == (37, 0)
    # batched, ragged
    S, N, M = 5, 16, 24
    A, B_ = np.stack([_boxes(g, N) for _ in range(S)]), np.stack([_boxes(g, M) for _ in range(S)])
    na, nb = np.array([16, 0, 7, 1, 12], np.int32), np.array([24, 5, 0, 24, 3], np.int32)
    cost = torch.full((S, N, M), -7.0, device="cuda")
    lib = _lib.load()
    dA, dB, dna, dnb = torch.as_tensor(A).cuda(), torch.as_tensor(B_).cuda(), torch.as_tensor(na).cuda(), torch.as_tensor(nb).cuda()
    _lib.check(lib.b2_iou_cost(_lib.ptr(dA), _lib.ptr(dB), None, _lib.ptr(dna), _lib.ptr(dnb), S, N, M, _lib.ptr(cost), _lib.stream_ptr()))
    cost = cost.cpu().numpy()
    for s in range(S):
        assert np.array_equal(cost[s, :na[s], :nb[s]], 1 - obt.bbox_iou_f32(A[s, :na[s]], B_[s, :nb[s]]).reshape(na[s], nb[s]))
        assert np.all(cost[s, na[s]:] == -7.0) and np.all(cost[s, :, nb[s]:] == -7.0)


def _check_assignment(c, th, x, y):
    n, m = c.shape
    _, xo, yo = obt.lapjv(c, extend_cost=True, cost_limit=th)
    assert len({j for j in x if j >= 0}) == int((x >= 0).sum())                  # a matching
    for i, j in enumerate(x):
        assert j < 0 or (y[j] == i and c[i, j] <= th)
    assert int((y >= 0).sum()) == int((x >= 0).sum())
    tot = lambda xx: sum(float(c[i, j]) - th for i, j in enumerate(xx) if j >= 0)
    assert abs(tot(x) - tot(xo)) < 1e-9, (tot(x), tot(xo))                       # the optimum ...
    assert np.array_equal(x, xo) and np.array_equal(y, yo)                       # ... and, without exact ties, the same matching


def test_linear_assignment_matches_reference_cases_and_oracle():
    """b2_linear_assignment against (a) the cases solved through the reference's matching.linear_assignment (golden) and (b) the
    restated lapjv on random rectangular problems: tall, wide, 1 x m, sizes up to 300 x 300 (max_det), IoU-like sparse costs where
    most entries sit above the limit, and a ragged batch."""
    import torch

    from b200dt import byte_tracker as bt

    gold = np.load(os.path.join(G, "bytetrack.npz"))
    for k in range(5):
        c, x, th = gold[f"lap{k}_cost"], gold[f"lap{k}_x"], float(gold[f"lap{k}_thresh"])
        m_, ua, ub = bt.linear_assignment(c, th)
        xx = np.full(len(x), -1)
        for i, j in m_:
            xx[i] = j
        assert np.array_equal(xx, x)
        assert sorted(ua) == [i for i in range(len(x)) if x[i] < 0] and sorted(ub) == sorted(set(range(c.shape[1])) - set(x[x >= 0]))
    g = np.random.default_rng(1)
    for n, m, th in [(1, 1, 0.5), (1, 9, 0.8), (9, 1, 0.8), (17, 40, 0.8), (40, 17, 0.6), (64, 64, 0.9), (120, 150, 0.8), (300, 300, 0.8)]:
        c = g.uniform(0, 1, (n, m)).astype(np.float32)
        x, y = bt.linear_assignment_device(torch.as_tensor(c).cuda()[None], th)
        _check_assignment(c, th, x[0].cpu().numpy(), y[0].cpu().numpy())
    # IoU-like: boxes of one scene against jittered copies + clutter (most costs are exactly 1.0 = no overlap)
    a = _boxes(g, 80, 600)
    b = np.concatenate([a[:60] + g.normal(0, 1.5, (60, 4)).astype(np.float32), _boxes(g, 25, 600)])
    c = (1 - obt.bbox_iou_f32(a, b)).astype(np.float32)
    x, y = bt.linear_assignment_device(torch.as_tensor(c).cuda()[None], 0.8)
    _check_assignment(c, 0.8, x[0].cpu().numpy(), y[0].cpu().numpy())
    assert (x[0] >= 0).sum() >= 50
    # empty problems and a ragged batch in one launch
    assert bt.linear_assignment(np.zeros((0, 4), np.float32), 0.8)[0].shape == (0, 2)
    S, N, M = 6, 20, 30
    C = g.uniform(0, 1, (S, N, M)).astype(np.float32)
    na, nb = np.array([20, 0, 5, 20, 1, 13], np.int32), np.array([30, 7, 0, 2, 30, 13], np.int32)
    x, y = bt.linear_assignment_device(torch.as_tensor(C).cuda(), 0.7, torch.as_tensor(na).cuda(), torch.as_tensor(nb).cuda())
    x, y = x.cpu().numpy(), y.cpu().numpy()
    for s in range(S):
        assert np.all(x[s, na[s]:] == -1) and np.all(y[s, nb[s]:] == -1)
        if na[s] and nb[s]:
            _check_assignment(C[s, :na[s], :nb[s]], 0.7, x[s, :na[s]], y[s, :nb[s]])
        else:
            assert np.all(x[s] == -1) and np.all(y[s] == -1)


@pytest.mark.parametrize("tag", ["default", "nofuse"])
def test_bytetracker_matches_reference(tag):
    """b200dt.byte_tracker.BYTETracker.update over the scripted 90-frame scene against the reference's own rows
    (tests/golden/bytetrack.npz): the same tracks every frame -- id, score, class, detection index bit-exact; boxes within 2e-3 px
    (fp32 Kalman state on the device, float64 numpy in the reference; the a18 gate is 1e-5 relative to coordinates of ~600)."""
    from b200dt import byte_tracker as bt
    from b200dt.predictor import Boxes

    g = np.load(os.path.join(G, "bytetrack.npz"))
    trk = bt.BYTETracker(dict(BYTE_PARAMS[tag]) or None)
    off = so = 0
    worst = 0.0
    for f, d in enumerate(bytetrack_script()):
        r = np.asarray(trk.update(Boxes(d, (512, 640))), dtype=np.float32).reshape(-1, 8)
        n = int(g[f"{tag}_counts"][f])
        ref = g[f"{tag}_rows"][off:off + n]
        off += n
        assert len(r) == n, (f, len(r), n)
        assert np.array_equal(r[:, 4:], ref[:, 4:]), f
        if n:
            worst = max(worst, float(np.abs(r[:, :4] - ref[:, :4]).max()))
        ns = int(g[f"{tag}_nstate"][f])
        st = sorted(((q.track_id, q.state) for q in trk.tracked_stracks + trk.lost_stracks), key=lambda q: q[0])
        assert [q[0] for q in st] == list(g[f"{tag}_ids"][so:so + ns]) and [q[1] for q in st] == list(g[f"{tag}_state"][so:so + ns]), f
        so += ns
    assert worst < 2e-3, worst


def test_yolo_track_attaches_ids():
    """YOLO.track (engine/model.py:559-591 + trackers/track.py:72-102): a short synthetic video -> Results whose boxes carry ids
    (7 columns), stable for the seeded blobs across frames; persist=True continues the same tracker."""
    from b200dt import synth
    from b200dt.predictor import YOLO

    vid = synth.IRStream(seed=7, h=512, w=640)
    frames = [vid.frame() for _ in range(6)]
    model = YOLO("yolov8n-p2.yaml")
    res = model.track(frames[:4], conf=0.15, iou=0.6)
    assert len(res) == 4
    assert all(r.boxes.is_track and r.boxes.data.shape[1] == 7 for r in res if len(r))
    ids = [set(np.asarray(r.boxes.id).astype(int).tolist()) for r in res if len(r) and r.boxes.is_track]
    assert ids and len(ids[-1] & ids[-2]) >= max(1, len(ids[-1]) // 2)              # tracks persist from frame to frame
    nxt = model.track(frames[4:], conf=0.15, iou=0.6, persist=True)
    ids2 = set(np.asarray(nxt[0].boxes.id).astype(int).tolist())
    assert len(ids2 & ids[-1]) >= max(1, len(ids2) // 2)
    bot = model.track(frames[:3], conf=0.15, iou=0.6, tracker="botsort.yaml")               # BoT-SORT: XYWH filter + GMC on the frames
    assert all(r.boxes.is_track for r in bot if len(r))
    with pytest.raises(AssertionError):
        model.track(frames[:1], tracker="deepsort.yaml")


@pytest.mark.parametrize("tag,use_img", [("botsort", True), ("botsort_nogmc", False)])
def test_botsort_matches_reference(tag, use_img):
    """b200dt.byte_tracker.BOTSORT (no ReID) against the reference's BOTSORT.update on a shaking-camera scene: with the frames
    (sparse-optical-flow GMC, OpenCV on the host as in the reference) and without (img=None): same ids / scores / classes /
    indices every frame, boxes within 2e-3 px."""
    from b200dt import byte_tracker as bt
    from b200dt.predictor import Boxes

    from golden_common import botsort_scene

    g = np.load(os.path.join(G, "bytetrack.npz"))
    frames, dets = botsort_scene()
    trk = bt.BOTSORT()
    off, worst = 0, 0.0
    for f, d in enumerate(dets):
        r = np.asarray(trk.update(Boxes(d, frames[f].shape[:2]), frames[f] if use_img else None), dtype=np.float32).reshape(-1, 8)
        n = int(g[f"{tag}_counts"][f])
        ref = g[f"{tag}_rows"][off:off + n]
        off += n
        assert len(r) == n, (f, len(r), n)
        assert np.array_equal(r[:, 4:], ref[:, 4:]), f
        if n:
            worst = max(worst, float(np.abs(r[:, :4] - ref[:, :4]).max()))
    assert worst < 2e-3, worst


def test_bytetracker_edge_frames_match_oracle():
    """Frames the scripted scene does not have: no detections at all (every track goes lost, then is removed after track_buffer
    frames), a single detection, only low-score detections, and recovery afterwards -- against the restated tracker fed the same
    boxes (ids / states bit-exact, boxes within the fp32 tolerance)."""
    from b200dt import byte_tracker as bt
    from b200dt.predictor import Boxes

    g = np.random.default_rng(4)
    base = _boxes(g, 6, 400)
    seq = []
    for f in range(70):
        if 10 <= f < 14 or 30 <= f < 65:
            d = np.zeros((0, 6), np.float32)                                  # nothing detected
        else:
            b = base + np.float32(f) * np.array([1.5, 0.5, 1.5, 0.5], np.float32)
            conf = np.full(6, 0.8, np.float32)
            if f == 20:
                b, conf = b[:1], conf[:1]                                     # a single box
            if f in (22, 23):
                conf = np.full(len(b), 0.15, np.float32)                      # low-score boxes only: second association alone
            d = np.concatenate([b, conf[:, None], np.zeros((len(b), 1), np.float32)], 1)
        seq.append(d)
    trk, ora = bt.BYTETracker(None), obt.BYTETracker()
    xywh = lambda d: np.stack([(d[:, 0] + d[:, 2]) / 2, (d[:, 1] + d[:, 3]) / 2, d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]], 1).astype(np.float32)
    for f, d in enumerate(seq):
        r = np.asarray(trk.update(Boxes(d, (512, 640))), dtype=np.float32).reshape(-1, 8)
        o = ora.update(xywh(d), d[:, 4], d[:, 5]).reshape(-1, 8)
        assert r.shape == o.shape, (f, r.shape, o.shape)
        assert np.array_equal(r[:, 4:], o[:, 4:]), f
        if len(r):
            assert np.abs(r[:, :4] - o[:, :4]).max() < 2e-3, f
        sa = sorted((t.track_id, t.state) for t in trk.tracked_stracks + trk.lost_stracks)
        sb = sorted((t.track_id, t.state) for t in ora.tracked_stracks + ora.lost_stracks)
        assert sa == sb, f
    # the six original tracks timed out in the long gap (track_buffer = 30 frames); the boxes of the last frames founded new ones
    assert len(trk.removed_stracks) >= 6 and all(t.track_id > 6 for t in trk.tracked_stracks + trk.lost_stracks)
