"""GPU parity of the detect path (engine forward, DFL decode, NMS, predict) against the oracle and the
reference-generated golden vectors."""
import os

import numpy as np
import pytest

import b200dt  # noqa: F401
from b200dt import cfg, synth, weights
from oracle import net as onet
from oracle import postprocess as pp

from golden_common import synth_pred

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _torch():
    import torch

    return torch


def _rel_l2(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


@pytest.fixture(scope="module")
def n_p2():
    spec = cfg.resolve("yolov8n-p2")
    return spec, onet.build_spec("yolov8n-p2"), weights.synthetic_state_dict(spec, seed=0)


def test_engine_layers_match_bf16_oracle(n_p2):
    """Every module output of the CUDA engine vs the oracle evaluated with the engine's rounding points
    (bf16 weights/activations, fp32 accumulate).  Per-layer error budget: a few bf16 roundings."""
    from b200dt.engine import Engine

    torch = _torch()
    spec, ospec, sd = n_p2
    B, H, W = 2, 64, 96
    x = np.random.default_rng(0).random((B, 3, H, W), dtype=np.float32)
    eng = Engine(spec, sd, B, H, W)
    eng.forward_tensor(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    net = onet.Net(ospec, sd, "bf16")
    heads = net.forward(x, record=True)
    worst = 0.0
    for name, ref in net.trace.items():
        if name.startswith("layer."):
            continue
        got = eng.activation(name).cpu().numpy()
        assert got.shape == ref.shape, name
        err = _rel_l2(got, ref)
        worst = max(worst, err)
        assert err < 2e-2, (name, err)
    # head logits: [B][h*w][64+nc]
    for l, hd in enumerate(heads):
        got = eng.level_logits(l).float().cpu().numpy()[..., :64 + spec["nc"]]
        ref = hd.reshape(B, hd.shape[1], -1).transpose(0, 2, 1)
        assert _rel_l2(got, ref) < 2e-2, (l, _rel_l2(got, ref))
    # CUDA graph replay gives the same bits as the first (capturing) run
    first = [eng.level_logits(l).clone() for l in range(eng.n_levels)]
    eng.forward_tensor(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    for l in range(eng.n_levels):
        assert torch.equal(first[l], eng.level_logits(l))
    eng.use_graph(False)
    eng.forward_tensor(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    for l in range(eng.n_levels):
        assert torch.equal(first[l], eng.level_logits(l))


def test_engine_vs_reference_fp32_golden(n_p2):
    """bf16 engine vs the reference's own fp32 forward (tests/golden/net_n_p2_small.npz): decoded boxes within
    1e-2 relative of the box scale for the bulk of anchors (H1: random nets amplify bf16 rounding)."""
    from b200dt import ops
    from b200dt.engine import Engine

    torch = _torch()
    spec, _, sd = n_p2
    g = np.load(os.path.join(G, "net_n_p2_small.npz"))
    x = np.random.default_rng(0).random((1, 3, 64, 96), dtype=np.float32)
    eng = Engine(spec, sd, 1, 64, 96)
    eng.forward_tensor(torch.from_numpy(x).cuda())
    post = ops.DetectPost(1, eng.level_h, eng.level_w, eng.level_stride, eng.nc, eng.lstride)
    dense = torch.zeros((1, 4 + eng.nc, eng.num_anchors), dtype=torch.float32, device="cuda")
    post.decode(eng.level_ptrs, 0.15, dense_out=dense)
    y = dense.cpu().numpy()
    ref = g["y"]
    box_err = np.abs(y[:, :4] - ref[:, :4]).max(1) / np.maximum(ref[:, 2:4].max(1), 1.0)
    assert np.median(box_err) < 1e-2 and np.quantile(box_err, 0.99) < 8e-2, (np.median(box_err), np.quantile(box_err, 0.99))
    assert np.abs(y[:, 4:] - ref[:, 4:]).max() < 0.2


def test_decode_matches_oracle_on_same_logits():
    """DFL + anchors + sigmoid + candidate filter on given bf16 logits: exact arithmetic restatement."""
    from b200dt import ops

    torch = _torch()
    g = np.random.default_rng(1)
    B, nc = 3, 80
    hs, ws, st = [16, 8, 4, 2], [24, 12, 6, 3], [4, 8, 16, 32]
    lstride = 144
    maps, bufs = [], []
    for h, w in zip(hs, ws):
        m = onet.bf16_round((g.standard_normal((B, 64 + nc, h, w)) * 2.0).astype(np.float32))
        m[:, 64:] -= 3.0
        m = onet.bf16_round(m)
        maps.append(m)
        bufs.append(torch.from_numpy(np.ascontiguousarray(m.reshape(B, 64 + nc, h * w).transpose(0, 2, 1))).cuda().to(torch.bfloat16).contiguous())
    ref = pp.decode(maps, st, nc)                                       # (B, 84, A)
    post = ops.DetectPost(B, hs, ws, st, nc, lstride)
    dense = torch.zeros((B, 4 + nc, post.A), dtype=torch.float32, device="cuda")
    conf = 0.15
    post.decode(bufs, conf, dense_out=dense)
    y = dense.cpu().numpy()
    np.testing.assert_allclose(y[:, :4], ref[:, :4], rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(y[:, 4:], ref[:, 4:], rtol=2e-5, atol=1e-6)
    # candidate set == anchors whose best score > conf, outside a tiny band around the threshold
    cnt = post.cand_count.cpu().numpy()
    cand = post.cand.cpu().numpy()
    cidx = post.cand_idx.cpu().numpy()
    for b in range(B):
        best = ref[b, 4:].max(0)
        sure = set(np.nonzero(best > conf + 1e-5)[0].tolist())
        maybe = set(np.nonzero(best > conf - 1e-5)[0].tolist())
        got = set(cidx[b, :cnt[b]].tolist())
        assert sure <= got <= maybe
        order = np.argsort(cidx[b, :cnt[b]])
        rows = cand[b, :cnt[b]][order]
        a = cidx[b, :cnt[b]][order]
        xyxy = pp.xywh2xyxy(ref[b, :4].T[a])
        np.testing.assert_allclose(rows[:, :4], xyxy, rtol=2e-5, atol=3e-4)
        np.testing.assert_array_equal(rows[:, 5].astype(int), ref[b, 4:].argmax(0)[a])


@pytest.mark.parametrize("ci", range(6))
@pytest.mark.parametrize("mode", ["exact", "legacy"])
def test_nms_bit_exact_vs_reference(ci, mode):
    """non_max_suppression (candidate filter, class offsets, greedy NMS, max_det) vs the reference's outputs."""
    from b200dt import ops

    torch = _torch()
    g = np.load(os.path.join(G, "nms_cases.npz"))
    seed, B, nc, A, conf, iou, max_det, agn = g[f"c{ci}_cfg"]
    pred = synth_pred(int(seed), int(B), int(nc), int(A), frac=0.05 if ci == 0 else 0.15)
    classes = g[f"c{ci}_classes"].tolist() if f"c{ci}_classes" in g.files else None
    out = ops.non_max_suppression(torch.from_numpy(pred).cuda(), float(conf), float(iou), classes=classes, agnostic=bool(agn),
                                  max_det=int(max_det), mode=mode)
    for b in range(int(B)):
        ref = g[f"c{ci}_{mode}_{b}"]
        got = out[b].cpu().numpy()
        assert got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)


def test_nms_edge_cases():
    from b200dt import ops

    torch = _torch()
    # no candidates at all
    pred = torch.zeros((2, 6, 50), device="cuda")
    out = ops.non_max_suppression(pred, 0.25, 0.45)
    assert [tuple(o.shape) for o in out] == [(0, 6), (0, 6)]
    # SURVEY 8c hand case: legacy early exit keeps all, exact suppresses the duplicate
    boxes = np.array([[200, 200, 210, 210], [0, 0, 10, 10], [0, 0, 10, 10.5]], np.float32)
    xywh = np.stack([(boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2, boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]], 0)
    pred = np.concatenate([xywh, np.array([[.9, .8, .7]], np.float32)], 0)[None]
    ex = ops.non_max_suppression(torch.from_numpy(pred).cuda(), 0.25, 0.6, mode="exact")[0]
    lg = ops.non_max_suppression(torch.from_numpy(pred).cuda(), 0.25, 0.6, mode="legacy")[0]
    assert len(ex) == 2 and len(lg) == 3
    with pytest.raises(AssertionError):
        ops.non_max_suppression(torch.from_numpy(pred).cuda(), 1.5, 0.6)
    # many candidates (global-memory sort path) against the oracle, bit-exact
    big = synth_pred(9, 1, 3, 6000, frac=0.3)
    got = ops.non_max_suppression(torch.from_numpy(big).cuda(), 0.05, 0.5, max_det=300)[0].cpu().numpy()
    ref = pp.non_max_suppression(big, 0.05, 0.5, max_det=300, mode="exact")[0]
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("tag,hw", [("512x640", (512, 640)), ("500x640", (500, 640))])
def test_predict_matches_reference_golden(tag, hw, n_p2):
    """YOLO(...).predict on uint8 frames vs the reference's predict() (fp32 CPU): same detection set up to the
    bf16 noise floor of a random-weight net -- most rows matched within 1e-2 relative box error."""
    from b200dt.predictor import YOLO

    g = np.load(os.path.join(G, "predict_n_p2.npz"))
    frames = [synth.IRStream(seed=7, h=hw[0], w=hw[1]).frame(), synth.IRStream(seed=8, h=hw[0], w=hw[1]).frame()]
    model = YOLO("yolov8n-p2.yaml")
    res = model.predict(frames, conf=0.15, iou=0.6, verbose=False)
    assert len(res) == 2
    for b, r in enumerate(res):
        ref = g[f"{tag}_exact_{b}"]
        d = r.boxes.data.cpu().numpy()
        assert r.orig_shape == hw and d.shape[1] == 6
        assert np.all(np.diff(d[:, 4]) <= 0)                              # sorted by descending confidence
        assert d[:, [0, 2]].min() >= 0 and d[:, [0, 2]].max() <= hw[1] and d[:, [1, 3]].max() <= hw[0]
        matched = 0
        for row in ref:
            c = d[d[:, 5] == row[5]]
            if len(c):
                scale = max(row[2] - row[0], row[3] - row[1], 1.0)
                e = np.abs(c[:, :4] - row[:4]).max(1) / scale
                if e.min() < 1e-2 * 4:
                    matched += 1
        assert matched >= 0.6 * len(ref), (matched, len(ref), len(d))
    # the reference API surface
    r = res[0]
    assert r.boxes.xyxy.shape[1] == 4 and r.boxes.conf.ndim == 1 and r.boxes.cls.ndim == 1 and r.boxes.id is None
    xy = r.boxes.xyxy.cpu().numpy()
    assert xy.dtype == np.float32


def test_predict_matches_bf16_oracle_end_to_end(n_p2):
    """Same frames through the oracle with the engine's rounding points: identical detection set except rows
    whose score or IoU sits within the bf16 noise band of a threshold."""
    from b200dt.predictor import YOLO

    spec, ospec, sd = n_p2
    frames = [synth.IRStream(seed=21, h=96, w=128).frame(), synth.IRStream(seed=22, h=96, w=128).frame()]
    model = YOLO("yolov8n-p2.yaml")
    res = model.predict(frames, conf=0.15, iou=0.6, imgsz=128)
    x = pp.preprocess(frames)
    heads = onet.Net(ospec, sd, "bf16").forward(x)
    y = pp.decode(heads, [4, 8, 16, 32], 80)
    ref = pp.non_max_suppression(y, 0.15, 0.6, mode="exact")
    for b in range(2):
        d = res[b].boxes.data.cpu().numpy()
        rb = ref[b].copy()
        rb[:, :4] = pp.scale_boxes((96, 128), rb[:, :4], (96, 128))
        matched = 0
        for row in rb:
            c = d[d[:, 5] == row[5]]
            if len(c) and (np.abs(c[:, :4] - row[:4]).max(1) + np.abs(c[:, 4] - row[4])).min() < 0.1:
                matched += 1
        assert matched >= 0.8 * len(rb) and abs(len(d) - len(rb)) <= max(3, 0.2 * len(rb)), (matched, len(rb), len(d))
