"""Helpers shared by tests/golden/make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import numpy as np

TRACK_KEYS = ("age", "hits", "hit_streak", "time_since_update", "lost_frames")


def pack_tracks(frames_out):
    """list (per frame) of list[dict] -> flat arrays."""
    rows = []
    traj_last, traj_len = [], []
    for f, tracks in enumerate(frames_out):
        for t in tracks:
            rows.append([f, int(t["track_id"][1:]), *np.asarray(t["bbox"], np.float64), float(t["confidence"]),
                         1.0 if t["status"] == "predicted" else 0.0, *[t[k] for k in TRACK_KEYS],
                         float(t["is_lost"]), *np.asarray(t["velocity"], np.float64), float(t["motion_confidence"]),
                         float(t["is_stable_motion"]), float(t["speed"]), float(t["direction"])])
            tr = np.asarray(t["trajectory"], np.float64).reshape(-1, 2)
            traj_len.append(len(tr))
            pad = np.zeros((30, 2)); pad[:len(tr)] = tr
            traj_last.append(pad)
    cols = ["frame", "id", "x1", "y1", "x2", "y2", "confidence", "predicted", *TRACK_KEYS, "is_lost",
            "vx", "vy", "motion_confidence", "is_stable_motion", "speed", "direction"]
    return np.asarray(rows, np.float64).reshape(-1, len(cols)), cols, np.asarray(traj_len), np.asarray(traj_last)



def synth_pred(seed, B, nc, A, frac=0.15, wh=(4, 60), span=600):
    """(B, 4+nc, A) xywh + class scores with clustered boxes so that suppression happens."""
    g = np.random.default_rng(seed)
    nclu = max(1, A // 6)
    centers = g.uniform(20, span, (B, nclu, 2))
    which = g.integers(0, nclu, (B, A))
    xy = np.take_along_axis(centers, which[..., None].repeat(2, -1), 1) + g.normal(0, 3, (B, A, 2))
    whs = g.uniform(*wh, (B, A, 2))
    sc = g.uniform(0, 1, (B, A, nc)) ** 6
    sc *= (g.random((B, A, 1)) < frac * 3)
    return np.concatenate([xy, whs, sc], 2).transpose(0, 2, 1).astype(np.float32)


