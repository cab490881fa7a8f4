"""The multi-stream detect+track step (pipeline.DetectTrackPipeline) through its three entry points: device frames on one
stream, device frames with NMS + tracker overlapped under the next forward, and pinned host frames with uploads / downloads on
their own streams.  All three must produce the same track rows, bit for bit, step after step, and the single-stream result
must equal YOLO.predict + EnhancedMultiTargetTracker.update called frame by frame (the reference driver's loop,
kalman/aircraft_detection_tracking.py:88-109)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

H, W = 96, 128
KW = dict(max_lost_frames=150, min_hits=1, iou_threshold=0.1)


def _frames(S, T):
    from b200dt import synth

    vids = [synth.IRStream(seed=40 + s, h=H, w=W, n_targets=4) for s in range(S)]
    return [np.stack([v.frame() for v in vids]) for _ in range(T)]


def _pipe(S, **kw):
    from b200dt.pipeline import DetectTrackPipeline

    return DetectTrackPipeline("yolov8n-p2", S, (H, W), 128, 0.15, 0.6, 300, capacity=1024, **KW, **kw)


def _snap(rows, counts):
    c = counts.cpu().numpy().copy()
    r = rows.cpu().numpy().copy()
    return [r[s, :c[s]].view(np.int32).copy() for s in range(len(c))]


def _sorted(block):
    return block[np.argsort(block[:, 0], kind="stable")]


def test_overlapped_and_host_steps_match_single_stream_step():
    S, T = 3, 6
    frames = _frames(S, T)
    plain, over, host = _pipe(S), _pipe(S, overlap_post=True), _pipe(S, overlap_post=True)
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    dev = [torch.from_numpy(f).cuda() for f in frames]
    ref = []
    for t in range(T):
        ref.append(_snap(*plain.step_device(dev[t])))
    torch.cuda.synchronize()
    # overlapped: results of step t are read after join(); the next step is already queued behind it on purpose
    for t in range(T):
        rows, counts = over.step_device(dev[t])
        over.join()
        got = _snap(rows, counts)
        for s in range(S):
            np.testing.assert_array_equal(_sorted(got[s]), _sorted(ref[t][s]))
    # back-to-back overlapped steps with no join in between: only the final state is compared
    over2 = _pipe(S, overlap_post=True)
    for t in range(T):
        rows, counts = over2.step_device(dev[t])
    over2.join()
    torch.cuda.synchronize()
    got = _snap(rows, counts)
    for s in range(S):
        np.testing.assert_array_equal(_sorted(got[s]), _sorted(ref[-1][s]))
    # host entry point (uploads and downloads on their own streams)
    for t in range(T):
        hr, hc = host.step_host(pinned[t])
    host.join()
    torch.cuda.synchronize()
    c = hc.numpy()
    for s in range(S):
        np.testing.assert_array_equal(_sorted(hr.numpy()[s, :c[s]].view(np.int32)), _sorted(ref[-1][s]))


def test_pipeline_step_equals_predict_plus_tracker_update_per_frame():
    from b200dt.predictor import YOLO
    from b200dt.tracker import EnhancedMultiTargetTracker, rows_to_dicts

    S, T = 2, 4
    frames = _frames(S, T)
    pipe = _pipe(S)
    model = YOLO("yolov8n-p2.yaml")
    trks = [EnhancedMultiTargetTracker(150, 1, 0.1, capacity=1024, max_dets=300) for _ in range(S)]
    for t in range(T):
        rows, counts = pipe.step_device(torch.from_numpy(frames[t]).cuda())
        c = counts.cpu().numpy()
        for s in range(S):
            res = model.predict(frames[t][s], conf=0.15, iou=0.6, imgsz=128)[0]
            d = res.boxes.data.cpu().numpy()
            want = trks[s].update([[*r[:4], r[4]] for r in d])
            got = rows_to_dicts(rows[s, :c[s]].cpu().numpy())
            assert [g["track_id"] for g in got] == [w["track_id"] for w in want]
            for g, w in zip(got, want):
                np.testing.assert_allclose(g["bbox"], w["bbox"], rtol=1e-5, atol=1e-3)


def test_yolo_loads_a_pt_checkpoint(tmp_path):
    """YOLO('best.pt') (engine/model.py:_load -> nn/tasks.py:1487-1521): a checkpoint pickled from classes of a package called
    `ultralytics` (a stand-in module tree built here from the synthetic state_dict: the reference package does not exist on the GPU
    box) is read without that package, and predicts exactly what YOLO(yaml, state_dict=...) predicts."""
    import sys
    import types

    import torch

    from b200dt import cfg, synth, weights
    from b200dt.predictor import YOLO

    spec = cfg.resolve("yolov8n-p2.yaml", nc=80)
    sd = weights.synthetic_state_dict(spec, seed=0)
    fake = types.ModuleType("ultralytics"); fake_nn = types.ModuleType("ultralytics.nn"); fake_tasks = types.ModuleType("ultralytics.nn.tasks")

    class Node(torch.nn.Module):
        pass

    class DetectionModel(torch.nn.Module):
        pass

    Node.__module__ = DetectionModel.__module__ = "ultralytics.nn.tasks"
    Node.__qualname__, DetectionModel.__qualname__ = "Node", "DetectionModel"
    fake_tasks.Node, fake_tasks.DetectionModel = Node, DetectionModel
    root = DetectionModel()
    for k, v in sd.items():
        *path, leaf = k.split(".")
        m = root
        for p in path:
            if p not in m._modules:
                m.add_module(p, Node())
            m = m._modules[p]
        t = torch.as_tensor(np.asarray(v))
        if leaf in ("running_mean", "running_var", "num_batches_tracked"):
            m.register_buffer(leaf, t)
        else:
            m.register_parameter(leaf, torch.nn.Parameter(t, requires_grad=False))
    root.yaml = dict(cfg.model_dict("yolov8n-p2.yaml"), nc=80)
    root.names = {i: f"c{i}" for i in range(80)}
    sys.modules.update({"ultralytics": fake, "ultralytics.nn": fake_nn, "ultralytics.nn.tasks": fake_tasks})
    try:
        path = str(tmp_path / "best.pt")
        torch.save({"epoch": 3, "model": None, "ema": root, "train_args": {}}, path)
    finally:
        for k in ("ultralytics", "ultralytics.nn", "ultralytics.nn.tasks"):
            sys.modules.pop(k, None)
    frames = [synth.IRStream(seed=60 + b, h=512, w=640).frame() for b in range(2)]
    a = YOLO(path).predict(frames, conf=0.15, iou=0.6)
    assert "ultralytics" not in sys.modules
    b = YOLO("yolov8n-p2.yaml", state_dict=sd, nc=80).predict(frames, conf=0.15, iou=0.6)
    assert a[0].names[5] == "c5"
    for ra, rb in zip(a, b):
        assert len(ra) > 0 and torch.equal(ra.boxes.data, rb.boxes.data)


def test_predict_and_track_from_files(tmp_path):
    """YOLO.predict / YOLO.track on file sources (data/loaders.py ingest in front of the engine): a directory of PNGs gives exactly
    the detections of the decoded arrays; a video file streams Results frame by frame with track ids and the file path."""
    import os
    import sys

    cv2 = pytest.importorskip("cv2")
    sys.path.insert(0, os.path.dirname(__file__))
    from test_host import _write_media

    from b200dt.predictor import YOLO

    imgs, clip = _write_media(tmp_path, n_img=3, n_vid=6, hw=(256, 320))
    model = YOLO("yolov8n-p2.yaml")
    a = model.predict(str(tmp_path / "im*.png"), conf=0.15, iou=0.6, batch=2)
    b = model.predict(imgs, conf=0.15, iou=0.6)
    assert len(a) == 3 and [os.path.basename(r.path) for r in a] == ["im0.png", "im1.png", "im2.png"]
    for ra, rb in zip(a, b):
        assert torch.equal(ra.boxes.data, rb.boxes.data)
    n = 0
    for r in model.track(clip, stream=True, conf=0.15, iou=0.6, batch=4):
        assert r.path.endswith("clip.avi") and r.orig_shape == (256, 320)
        assert len(r) == 0 or r.boxes.is_track
        n += 1
    assert n == 6


def test_track_over_a_streams_file(tmp_path):
    """YOLO.track(source='cams.streams'): LoadStreams hands one frame per source to one forward (batch = number of sources) and
    every source has its own tracker (trackers/track.py:62-68); Results come back source by source with ids attached."""
    cv2 = pytest.importorskip("cv2")
    from b200dt import synth
    from b200dt.predictor import YOLO

    clips = []
    for k in range(2):
        vid = synth.IRStream(seed=70 + k, h=256, w=320)
        path = str(tmp_path / f"cam{k}.avi")
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (320, 256))
        for _ in range(5):
            wr.write(vid.frame())
        wr.release()
        clips.append(path)
    (tmp_path / "cams.streams").write_text("\n".join(clips))
    model = YOLO("yolov8n-p2.yaml")
    out = list(model.track(str(tmp_path / "cams.streams"), stream=True, conf=0.15, iou=0.6, stream_buffer=True))
    assert len(out) == 10 and len(model.trackers) == 2 and model.dataset.mode == "stream"
    assert [r.path for r in out[:4]] == [model.dataset.sources[0], model.dataset.sources[1]] * 2
    assert model.trackers[0] is not model.trackers[1] and model.trackers[0].frame_id == 5 and model.trackers[1].frame_id == 5
    assert all(r.boxes.is_track for r in out if len(r))


def test_pipeline_run_over_stream_loader_equals_stepping(tmp_path):
    """DetectTrackPipeline.run(LoadStreams(...)): the multi-stream driver loop fed from files gives exactly the rows of stepping
    a second pipeline by hand on the same decoded frames."""
    cv2 = pytest.importorskip("cv2")
    from b200dt import synth
    from b200dt.loaders import LoadStreams

    clips = []
    for k in range(3):
        vid = synth.IRStream(seed=80 + k, h=H, w=W, n_targets=4)
        path = str(tmp_path / f"s{k}.avi")
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (W, H))
        for _ in range(6):
            wr.write(vid.frame())
        wr.release()
        clips.append(path)
    (tmp_path / "s.streams").write_text("\n".join(clips))
    a, b = _pipe(3), _pipe(3)
    got = list(a.run(LoadStreams(str(tmp_path / "s.streams"), buffer=True)))
    assert len(got) == 6
    ref_loader = LoadStreams(str(tmp_path / "s.streams"), buffer=True)
    for (src, rows, counts), (_, frames, _) in zip(got, ref_loader):
        r2, c2 = b.step_device(torch.from_numpy(np.stack(frames)).cuda())
        b.join(); torch.cuda.synchronize()
        c2 = c2.cpu().numpy()
        assert np.array_equal(counts, c2)
        for s in range(3):
            assert np.array_equal(rows[s, :counts[s]], r2[s, :c2[s]].cpu().numpy())
    assert sum(int(c.sum()) for _, _, c in got) > 0
