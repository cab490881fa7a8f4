"""Oracle (numpy, CPU) for letterbox, DFL decode, NMS and box rescale -- test infrastructure only.

Restates:
  letterbox geometry      ultralytics/data/augment.py:1692-1733 (LetterBox.__call__), pad value 114,
                          engine/predictor.py:152-175 (BGR->RGB, HWC->CHW, /255)
  cv2.resize INTER_LINEAR third-party (opencv-python>=4.6, pyproject.toml:66; call site augment.py:1718): uint8 path of
                          cv::resize -- 11-bit fixed-point coefficients, horizontal pass in int32, vertical pass
                          ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2; pinned against cv2 itself in
                          tests/test_oracle_vs_golden.py::test_resize_oracle_is_cv2
  Detect._inference       ultralytics/nn/modules/head.py:152-187
  DFL.forward             ultralytics/nn/modules/block.py:78-81
  make_anchors/dist2bbox  ultralytics/utils/tal.py:367-391
  non_max_suppression     ultralytics/utils/nms.py:59-167 (multi_label=False, rotated=False path)
  TorchNMS.nms            ultralytics/utils/nms.py:237-304   ("legacy" branch, with its early exit)
  torchvision.ops.nms     third-party (torchvision>=0.9, pyproject.toml:73; call site nms.py:155):
                          greedy, suppress iou > thr, stable descending-score order  ("exact" branch)
  xywh2xyxy               ultralytics/utils/ops.py:277-294
  scale_boxes/clip_boxes  ultralytics/utils/ops.py:105-138, :157-183
All float arithmetic is fp32 in the same operation order as the reference.
"""
from __future__ import annotations

import numpy as np

F = np.float32
MAX_WH = 7680      # nms.py:26
MAX_NMS = 30000    # nms.py:25


def letterbox_geometry(h0, w0, new_shape=(640, 640), auto=True, stride=32, scaleup=True, center=True):
    """Returns (r, new_unpad (w,h), top, bottom, left, right) exactly as LetterBox computes them."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return r, new_unpad, top, bottom, left, right


def letterbox_pad_only(img, new_shape=(640, 640), auto=True, stride=32):
    """Letterbox for frames that need no resize (r == 1): constant 114 border (copyMakeBorder)."""
    h0, w0 = img.shape[:2]
    r, new_unpad, top, bottom, left, right = letterbox_geometry(h0, w0, new_shape, auto, stride)
    assert (w0, h0) == new_unpad, "resize path not covered by letterbox_pad_only"
    out = np.full((h0 + top + bottom, w0 + left + right, img.shape[2]), 114, img.dtype)
    out[top:top + h0, left:left + w0] = img
    return out


def resize_bilinear_u8(img, dh, dw):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC, bit for bit."""
    sh, sw = img.shape[:2]
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)              # scale = 1 / inv_scale, as cv::resize forms it
    fx = ((np.arange(dw) + 0.5) * scale_x - 0.5).astype(F)            # double arithmetic, rounded to float once
    sx = np.floor(fx).astype(np.int64)
    fx = fx - sx.astype(F)
    lo, hi = sx < 0, sx >= sw - 1
    fx[lo | hi] = 0
    sx[lo], sx[hi] = 0, sw - 1
    fy = ((np.arange(dh) + 0.5) * scale_y - 0.5).astype(F)
    sy = np.floor(fy).astype(np.int64)
    fy = fy - sy.astype(F)
    a0, a1 = np.rint((F(1) - fx) * F(2048)).astype(np.int64), np.rint(fx * F(2048)).astype(np.int64)   # saturate_cast<short>(rint)
    b0, b1 = np.rint((F(1) - fy) * F(2048)).astype(np.int64), np.rint(fy * F(2048)).astype(np.int64)
    sx1, sy0, sy1 = np.minimum(sx + 1, sw - 1), np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    im = img.astype(np.int64)
    s0 = im[sy0][:, sx] * a0[None, :, None] + im[sy0][:, sx1] * a1[None, :, None]
    s1 = im[sy1][:, sx] * a0[None, :, None] + im[sy1][:, sx1] * a1[None, :, None]
    v = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def letterbox(img, new_shape=(640, 640), auto=True, stride=32):
    """LetterBox.__call__ (augment.py:1692-1733) in full: resize when the shape changes, then the 114 border."""
    h0, w0 = img.shape[:2]
    r, new_unpad, top, bottom, left, right = letterbox_geometry(h0, w0, new_shape, auto, stride)
    if (w0, h0) != new_unpad:
        img = resize_bilinear_u8(img, new_unpad[1], new_unpad[0])
    out = np.full((img.shape[0] + top + bottom, img.shape[1] + left + right, img.shape[2]), 114, img.dtype)
    out[top:top + img.shape[0], left:left + img.shape[1]] = img
    return out


def preprocess(imgs_bgr_u8):
    """predictor.preprocess for same-shape uint8 HWC BGR frames: -> (B,3,H,W) fp32 RGB in [0,1]."""
    im = np.stack(imgs_bgr_u8)[..., ::-1].transpose(0, 3, 1, 2)
    return np.ascontiguousarray(im).astype(F) / F(255)


def make_anchors(level_hw, strides, offset=0.5):
    pts, st = [], []
    for (h, w), s in zip(level_hw, strides):
        sx = np.arange(w, dtype=F) + F(offset)
        sy = np.arange(h, dtype=F) + F(offset)
        yy, xx = np.meshgrid(sy, sx, indexing="ij")
        pts.append(np.stack([xx, yy], -1).reshape(-1, 2))
        st.append(np.full((h * w, 1), s, F))
    return np.concatenate(pts), np.concatenate(st)


def decode(level_maps, strides, nc):
    """Detect._inference: list of (B, 64+nc, H, W) -> (B, 4+nc, A) [cx,cy,w,h, sigmoid(cls)...] fp32."""
    B = level_maps[0].shape[0]
    no = 64 + nc
    x_cat = np.concatenate([m.reshape(B, no, -1) for m in level_maps], 2).astype(F)
    anchors, st = make_anchors([m.shape[2:] for m in level_maps], strides)
    box, cls = x_cat[:, :64], x_cat[:, 64:]
    A = box.shape[2]
    b = box.reshape(B, 4, 16, A)
    b = b - b.max(2, keepdims=True)
    e = np.exp(b, dtype=F)
    p = e / e.sum(2, keepdims=True, dtype=F)
    dist = (p * np.arange(16, dtype=F)[None, None, :, None]).sum(2, dtype=F)      # (B,4,A) ltrb
    a = anchors.T[None]                                                           # (1,2,A)
    lt, rb = dist[:, :2], dist[:, 2:]
    x1y1, x2y2 = a - lt, a + rb
    c_xy = (x1y1 + x2y2) / F(2)
    wh = x2y2 - x1y1
    dbox = np.concatenate([c_xy, wh], 1) * st.T[None]
    return np.concatenate([dbox, (F(1) / (F(1) + np.exp(-cls, dtype=F))).astype(F)], 1).astype(F)


def xywh2xyxy(x):
    y = np.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / F(2)
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def _iou_one_to_many(b, rest):
    """IoU of box b (4,) against rest (n,4): inter / (area_a + area_b - inter), fp32, no eps."""
    xx1 = np.maximum(b[0], rest[:, 0]); yy1 = np.maximum(b[1], rest[:, 1])
    xx2 = np.minimum(b[2], rest[:, 2]); yy2 = np.minimum(b[3], rest[:, 3])
    w = np.clip(xx2 - xx1, 0, None).astype(F); h = np.clip(yy2 - yy1, 0, None).astype(F)
    inter = (w * h).astype(F)
    area_b = F((b[2] - b[0]) * (b[3] - b[1]))
    area_r = ((rest[:, 2] - rest[:, 0]) * (rest[:, 3] - rest[:, 1])).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter, (inter / ((area_b + area_r).astype(F) - inter)).astype(F)


def nms_exact(boxes, scores, thr):
    """torchvision.ops.nms semantics: visit in stable descending-score order, keep unless suppressed,
    suppress later boxes with iou > thr."""
    order = np.argsort(-scores, kind="stable")
    suppressed = np.zeros(len(order), bool)
    keep = []
    bs = boxes[order]
    for i in range(len(order)):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 < len(order):
            _, iou = _iou_one_to_many(bs[i], bs[i + 1:])
            suppressed[i + 1:] |= iou > F(thr)
    return np.asarray(keep, np.int64)


def nms_legacy(boxes, scores, thr):
    """TorchNMS.nms (nms.py:237-304) including the early exit at :290-296: when the current top box
    has zero intersection with every remaining box, ALL remaining boxes are kept and the loop ends."""
    if len(boxes) == 0:
        return np.zeros(0, np.int64)
    order = np.argsort(-scores, kind="stable")
    keep = []
    while len(order) > 0:
        i = order[0]
        keep.append(i)
        if len(order) == 1:
            break
        rest = order[1:]
        inter, iou = _iou_one_to_many(boxes[i], boxes[rest])
        if inter.sum(dtype=F) == 0:
            keep.extend(rest.tolist())
            break
        order = rest[iou <= F(thr)]
    return np.asarray(keep, np.int64)


def non_max_suppression(pred, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        max_det=300, nc=0, mode="exact", return_idx=False):
    """nms.py:59-167 for the detect path. pred: (B, 4+nc, A) xywh+scores. Returns list of (n,6) fp32.

    ``mode``: "exact" = torchvision.ops.nms branch (taken when torchvision is imported, nms.py:152-155),
    "legacy" = TorchNMS.nms branch (default for a bare ``predict(ndarray)``), see SURVEY.md H8.
    """
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    pred = np.asarray(pred, F)
    B = pred.shape[0]
    nc = nc or pred.shape[1] - 4
    out, idxs = [], []
    for xi in range(B):
        x = pred[xi].T.copy()                                    # (A, 4+nc)
        cand = x[:, 4:4 + nc].max(1) > F(conf_thres)
        ai = np.nonzero(cand)[0]
        x = x[cand]
        if not len(x):
            out.append(np.zeros((0, 6), F)); idxs.append(np.zeros(0, np.int64)); continue
        box = xywh2xyxy(x[:, :4])
        j = x[:, 4:4 + nc].argmax(1)
        conf = x[np.arange(len(x)), 4 + j]
        x = np.concatenate([box, conf[:, None], j[:, None].astype(F)], 1).astype(F)
        if classes is not None:
            m = np.isin(x[:, 5].astype(np.int64), np.asarray(classes))
            x, ai = x[m], ai[m]
        if not len(x):
            out.append(np.zeros((0, 6), F)); idxs.append(np.zeros(0, np.int64)); continue
        if len(x) > MAX_NMS:
            o = np.argsort(-x[:, 4], kind="stable")[:MAX_NMS]
            x, ai = x[o], ai[o]
        c = x[:, 5:6] * F(0 if agnostic else MAX_WH)
        boxes = (x[:, :4] + c).astype(F)
        keep = (nms_exact if mode == "exact" else nms_legacy)(boxes, x[:, 4], iou_thres)[:max_det]
        out.append(x[keep]); idxs.append(ai[keep])
    return (out, idxs) if return_idx else out


def scale_boxes(img1_shape, boxes, img0_shape):
    """ops.py:105-138 (ratio_pad=None, padding=True, xyxy) followed by clip_boxes :157-183. fp32 in place
    semantics on a copy."""
    boxes = np.array(boxes, F, copy=True)
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    boxes[..., 0] -= F(pad_x); boxes[..., 1] -= F(pad_y)
    boxes[..., 2] -= F(pad_x); boxes[..., 3] -= F(pad_y)
    boxes[..., :4] /= F(gain)
    h, w = img0_shape[:2]
    boxes[..., 0] = boxes[..., 0].clip(0, w); boxes[..., 1] = boxes[..., 1].clip(0, h)
    boxes[..., 2] = boxes[..., 2].clip(0, w); boxes[..., 3] = boxes[..., 3].clip(0, h)
    return boxes
