"""Full-size runs of the hot path (the configurations BASELINE.json / SURVEY.md 8d quote numbers on), checked through
size-independent properties instead of an oracle pass that would take hours:

  C4  256 streams x 512x640 frames, yolov8s-p2, detect + track:
        * a permutation of the streams permutes the results and changes nothing else (streams are independent,
          pipeline.shard_streams relies on it) -- bit-exact;
        * detections leave NMS in non-increasing score order, inside the frame, at most max_det per image;
        * NMS postcondition (utils/nms.py:129-160): no two kept boxes of one class overlap by more than iou_thres;
        * NMS is idempotent: feeding the kept boxes through NMS again keeps all of them.
  C3  track bank 256 streams x 4096 slots (1 048 576 tracks), 40 detections per stream and frame:
        * identical streams evolve identically (bit-exact);
        * a coasting track moves by exactly its velocity per predict() (constant-velocity model; the reference predicts twice
          on the frame a track is first missed);
        * the association is a matching: a detection updates at most one track, only tracks with IoU >= iou_threshold are
          updated, unmatched detections found new tracks.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

H, W = 512, 640
CONF, IOU = 0.15, 0.6


def _iou_matrix(b):
    x1 = torch.maximum(b[:, None, 0], b[None, :, 0]); y1 = torch.maximum(b[:, None, 1], b[None, :, 1])
    x2 = torch.minimum(b[:, None, 2], b[None, :, 2]); y2 = torch.minimum(b[:, None, 3], b[None, :, 3])
    inter = (x2 - x1).clamp(min=0) * (y2 - y1).clamp(min=0)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (area[:, None] + area[None, :] - inter)


def test_c4_full_size_step_properties():
    from b200dt import synth
    from b200dt.pipeline import DetectTrackPipeline

    S, T = 256, 3
    vids = [synth.IRStream(seed=500 + s, h=H, w=W) for s in range(8)]
    frames = []
    for _ in range(T):
        fr = [v.frame() for v in vids]
        frames.append(torch.from_numpy(np.stack([fr[s % 8] for s in range(S)])).cuda())
    kw = dict(capacity=2048, max_lost_frames=150, min_hits=1, iou_threshold=0.1)
    a = DetectTrackPipeline("yolov8s-p2", S, (H, W), 640, CONF, IOU, 300, **kw)
    b = DetectTrackPipeline("yolov8s-p2", S, (H, W), 640, CONF, IOU, 300, **kw)
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(3)).cuda()
    for t in range(T):
        ra, ca = a.step_device(frames[t])
        rb, cb = b.step_device(frames[t][perm].contiguous())
        torch.cuda.synchronize()
        # streams are independent: permuting the inputs permutes the outputs, bit for bit
        assert torch.equal(ca[perm], cb)
        ia, ib = ra.view(torch.int32)[perm], rb.view(torch.int32)
        live = torch.arange(ra.shape[1], device="cuda")[None, :] < cb[:, None]
        assert torch.equal(ia[live], ib[live])
        # detections of this step (post.out is what the tracker consumed)
        post = a.detect.post
        dets, cnt = post.out, post.out_count
        assert int(cnt.max()) <= 300 and int(cnt.min()) >= 0
        valid = torch.arange(dets.shape[1], device="cuda")[None, :] < cnt[:, None]
        sc = torch.where(valid, dets[..., 4], torch.full_like(dets[..., 4], -1.0))
        assert bool((sc[:, :-1] >= sc[:, 1:]).all()), "detections must leave NMS in score order"
        bx = dets[..., :4][valid]
        assert bool((bx[:, 0] >= 0).all() and (bx[:, 1] >= 0).all() and (bx[:, 2] <= W).all() and (bx[:, 3] <= H).all())
        assert bool((dets[..., 4][valid] > CONF).all())
    # NMS postcondition + idempotence on the last step's candidates (canvas coordinates: no scaling)
    post = a.detect.post
    d0, c0 = post.nms(IOU, scale=None)
    d0, c0 = d0.clone(), c0.clone()
    for s in (0, 1, 17, 100, 255):
        k = int(c0[s])
        box, cls = d0[s, :k, :4], d0[s, :k, 5]
        off = box + cls[:, None] * 7680.0                               # the reference's class offset (nms.py:144)
        iou = _iou_matrix(off)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= IOU + 1e-6, f"stream {s}: kept boxes overlap by {float(iou.max())}"
    post.cand[:, :300] = d0
    post.cand_idx[:, :300] = torch.arange(300, device="cuda", dtype=torch.int32)[None]
    post.cand_count.copy_(c0)
    d1, c1 = post.nms(IOU, scale=None)
    assert torch.equal(c1, c0)
    valid = torch.arange(300, device="cuda")[None, :] < c0[:, None]
    assert torch.equal(d1[valid], d0[valid]), "NMS of its own output must keep everything, in the same order"


def test_c3_full_size_bank_properties():
    from b200dt.tracker import TrackerBank

    S, C, D, THR = 256, 4096, 1024, 0.1
    bank = TrackerBank(S, C, D, 150, 1, THR)
    # fill every slot: 4096 well-separated 6x6 boxes on a 10 px grid, the same in every stream
    for r in range(C // D):
        idx = torch.arange(D, device="cuda") + r * D
        x, y = (idx % 64).float() * 10.0, (idx // 64).float() * 10.0
        boxes = torch.stack([x, y, x + 6, y + 6], 1)[None].repeat(S, 1, 1).contiguous()
        rows, counts = bank.update(boxes, torch.full((S,), D, dtype=torch.int32, device="cuda"), with_trajectory=False)
    torch.cuda.synchronize()
    assert bool((counts == C).all()), "every detection must have founded a track (1 048 576 live tracks)"
    rows0 = rows.clone()
    ids0 = rows0.view(torch.int32)[..., 0]
    assert torch.equal(rows0[0].expand_as(rows0), rows0), "identical streams must evolve identically"

    # frame A: 40 detections on top of existing tracks, shifted by (1, 0): they must update exactly those tracks
    g = torch.Generator(device="cuda").manual_seed(0)
    pick = torch.stack([torch.randperm(C, device="cuda", generator=g)[:40] for _ in range(4)])[torch.arange(S, device="cuda") % 4]
    px, py = (pick % 64).float() * 10.0 + 1.0, (pick // 64).float() * 10.0
    dets = torch.zeros((S, D, 6), device="cuda")
    dets[:, :40, 0], dets[:, :40, 1], dets[:, :40, 2], dets[:, :40, 3], dets[:, :40, 4] = px, py, px + 6, py + 6, 0.9
    cnt = torch.full((S,), 40, dtype=torch.int32, device="cuda")
    rows, counts = bank.update(dets, cnt, with_trajectory=False)
    torch.cuda.synchronize()
    rows1 = rows.clone()
    assert bool((counts == C).all()), "a matched detection must not found a new track"
    ir = rows1.view(torch.int32)
    tsu = ir[..., 10]                                                    # time_since_update
    assert bool(((tsu == 0).sum(1) == 40).all()), "exactly one track per detection was updated"
    # the updated tracks are the ones sitting under the detections: grid cell -> id from the fill (row order is not assumed)
    r0 = rows0[0]
    cell = (torch.round(r0[:, 1] / 10.0) + 64 * torch.round(r0[:, 2] / 10.0)).long()
    id_of_cell = torch.zeros(C, dtype=torch.int32, device="cuda")
    id_of_cell[cell] = ids0[0]
    for s in (0, 1, 2, 3, 255):
        upd = set(ir[s, :, 0][tsu[s] == 0].tolist())
        want = set(id_of_cell[pick[s]].tolist())
        assert upd == want
    # same inputs in streams s and s + 4 -> same outputs
    assert torch.equal(rows1[:4].repeat(S // 4, 1, 1), rows1)

    # frames B..: nothing detected -> every track coasts: predict is x <- F x with F = [[I, I], [0, I]] (kalman_tracker.py
    # predict), i.e. the centre and size move by exactly the state velocity and the velocity stays put
    zero = torch.zeros((S,), dtype=torch.int32, device="cuda")
    for s in (0, 129):
        x0, _, m0, _ = bank.export(s)
        assert len(x0) == C
    x0, _, m0, _ = bank.export(0)
    assert np.abs(x0[:, 4]).max() > 0.05, "the matched tracks must have picked up a velocity"
    for _ in range(3):
        rows, counts = bank.update(dets, zero, with_trajectory=False)
        torch.cuda.synchronize()
        assert bool((counts == C).all())
        x1, _, m1, _ = bank.export(0)
        np.testing.assert_array_equal(m1[:, 0], m0[:, 0])                 # same ids, same order
        # predict() runs once per frame -- twice on the frame a track is first missed (the reference's get_track_info ->
        # get_lost_prediction -> enhanced_long_term_predict(1), enhanced_aircraft_kalman_tracker.py:216-217, :319-333)
        k = m1[:, 4] - m0[:, 4]
        np.testing.assert_array_equal(k, np.where(m0[:, 4] == 0, 2, 1))
        np.testing.assert_allclose(x1[:, :4], x0[:, :4] + k[:, None] * x0[:, 4:], rtol=1e-6, atol=1e-4)
        np.testing.assert_allclose(x1[:, 4:], x0[:, 4:], rtol=1e-6, atol=1e-6)
        x0, m0 = x1, m1
    bank.close()


def _same_tracks(got, ref, where, box_rtol=1e-5):
    assert [t["track_id"] for t in got] == [t["track_id"] for t in ref], where
    for ta, tb in zip(got, ref):
        key = ("status", "age", "hits", "hit_streak", "time_since_update", "lost_frames", "is_lost")
        assert tuple(ta[k] for k in key) == tuple(tb[k] for k in key), (where, ta["track_id"])
        scale = max(1.0, float(np.abs(tb["bbox"]).max()))
        assert np.abs(np.asarray(ta["bbox"]) - np.asarray(tb["bbox"])).max() / scale < box_rtol, (where, ta["track_id"])


def test_c3_sized_stream_matches_oracle():
    """One stream of the C3 shape -- a 4096-slot bank holding 3072 tracks, 40 detections per frame of which 70 % come from
    existing tracks (+N(0,1) px) and 30 % are clutter -- against the float64 oracle, frame by frame: ids and
    lifecycle bit-exact, boxes to 1e-5.  Four streams run together (16 sweep chunks each); streams 0 and 3 are checked."""
    from b200dt.tracker import TrackerBank, rows_to_dicts
    from oracle import tracker as otr

    S, C, D, params = 4, 4096, 1024, (150, 1, 0.1)
    bank = TrackerBank(S, C, D, *params)
    g = np.random.default_rng(11)
    oracles = {0: otr.MultiTracker(*params), 3: otr.MultiTracker(*params)}
    frames = []
    for r in range(3):                                       # 3 x 1024 founding detections on a 10 px grid (6x6 boxes)
        idx = np.arange(D) + r * D
        x, y = (idx % 64) * 10.0, (idx // 64) * 10.0
        frames.append(np.stack([x, y, x + 6, y + 6, np.full(D, 0.9)], 1).astype(np.float32))
    for _ in range(8):
        pick = g.permutation(3 * D)[:40]
        x, y = (pick % 64) * 10.0, (pick // 64) * 10.0
        clutter = g.random(40) < 0.3
        x = np.where(clutter, g.uniform(0, 634, 40), x + g.normal(0, 1, 40))
        y = np.where(clutter, g.uniform(0, 474, 40), y + g.normal(0, 1, 40))
        frames.append(np.stack([x, y, x + 6, y + 6, np.full(40, 0.8)], 1).astype(np.float32))
    for f, d in enumerate(frames):
        dets = torch.zeros((S, D, 5), dtype=torch.float32)
        dets[:, :len(d)] = torch.from_numpy(d)[None]
        rows, counts = bank.update(dets.cuda(), torch.full((S,), len(d), dtype=torch.int32).cuda(), with_trajectory=False)
        c = counts.cpu().numpy()
        for s, o in oracles.items():
            _same_tracks(rows_to_dicts(rows[s, :c[s]].cpu().numpy()), o.update([r for r in d]), (s, f))
    for s, o in oracles.items():
        assert o.min_competing_gap > 1e-5 and o.min_thr_gap > 1e-6, (o.min_competing_gap, o.min_thr_gap)
        assert int(bank.export(s)[3][7]) == 0
    bank.close()


def test_c4_pipeline_tracker_matches_oracle_on_its_own_detections():
    """The headline workload's tracker stage (256 streams, capacity 2048, the synthetic IR frames bench.py uses): the NMS
    output the GPU itself produced for streams 0, 17 and 255 is fed, frame by frame, to the float64 oracle; ids / lifecycle
    must agree bit for bit and boxes to 1e-5 (kalman/enhanced_multi_target_tracker.py:42-132)."""
    from b200dt import synth
    from b200dt.pipeline import DetectTrackPipeline
    from b200dt.tracker import rows_to_dicts
    from oracle import tracker as otr

    S, T = 256, 14
    params = dict(max_lost_frames=150, min_hits=1, iou_threshold=0.1)
    vids = [synth.IRStream(seed=1000 + s, h=H, w=W) for s in range(32)]
    pipe = DetectTrackPipeline("yolov8s-p2", S, (H, W), 640, CONF, IOU, 300, capacity=2048, **params)
    sample = (0, 17, 255)
    oracles = {s: otr.MultiTracker(150, 1, 0.1) for s in sample}
    for t in range(T):
        fr = [v.frame() for v in vids]
        rows, counts = pipe.step_device(torch.from_numpy(np.stack([fr[s % 32] for s in range(S)])).cuda())
        torch.cuda.synchronize()
        dets, cnt = pipe.detect.post.out, pipe.detect.post.out_count
        c = counts.cpu().numpy()
        for s, o in oracles.items():
            d = dets[s, :int(cnt[s])].cpu().numpy()
            ref = o.update([np.asarray(r[:5], np.float32) for r in d])
            _same_tracks(rows_to_dicts(rows[s, :c[s]].cpu().numpy()), ref, (s, t))
    st = pipe.bank.stats_async().cpu().numpy()
    assert int(st[:, 5].sum()) == 0, "no detection may be dropped"
    for s, o in oracles.items():
        assert [int(v) for v in st[s, :5]] == [o.stats[k] for k in ("total_tracks_created", "total_tracks_terminated",
                                                                     "current_active_tracks", "long_term_predictions",
                                                                     "successful_recoveries")]


@pytest.mark.parametrize("name,B,HW", [("yolov8s-p2", 64, (640, 640)), ("yolov8x-p2", 32, (1280, 1280))], ids=["C2", "C5"])
def test_c2_c5_full_size_forward_decode_nms_properties(name, B, HW):
    """C2 / C5: BCHW tensors in [0,1) -> forward + decode + NMS at the quoted batch and resolution.
    Properties: bit-identical results when the same batch runs twice and when the batch is reversed (no cross-image
    coupling at any tile boundary), finite boxes in score order, NMS postcondition on a sample of images."""
    from b200dt import cfg, weights
    from b200dt.predictor import DetectPipeline

    spec = cfg.resolve(name, nc=80)
    pipe = DetectPipeline(spec, weights.synthetic_state_dict(spec, seed=0), B, HW[0], HW[1], 300)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((B, 3, HW[0], HW[1]), device="cuda", generator=g).to(torch.bfloat16)
    d0, c0 = pipe.run_tensor(x, CONF, IOU)
    d0, c0 = d0.clone(), c0.clone()
    d1, c1 = pipe.run_tensor(x, CONF, IOU)
    assert torch.equal(c0, c1) and torch.equal(d0, d1)
    d2, c2 = pipe.run_tensor(x.flip(0).contiguous(), CONF, IOU)
    torch.cuda.synchronize()
    assert torch.equal(c2.flip(0), c0)
    valid = torch.arange(300, device="cuda")[None, :] < c0[:, None]
    assert torch.equal(d2.flip(0)[valid], d0[valid])
    assert bool(torch.isfinite(d0[valid]).all())
    sc = torch.where(valid, d0[..., 4], torch.full_like(d0[..., 4], -1.0))
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    # postcondition on the unclipped boxes NMS itself saw (predict() clips to the frame afterwards, ops.scale_boxes)
    pipe.run_tensor(x, CONF, IOU)
    du, cu = pipe.post.nms(IOU, scale=None)
    assert torch.equal(cu, c0)
    for s in (0, B // 2, B - 1):
        k = int(cu[s])
        if k < 2:
            continue
        off = du[s, :k, :4] + du[s, :k, 5:6] * 7680.0
        iou = torch.nan_to_num(_iou_matrix(off), nan=0.0)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= IOU + 1e-6
