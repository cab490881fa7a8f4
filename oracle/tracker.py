"""Oracle (numpy float64, CPU) for the project's Kalman multi-target tracker -- test infrastructure only.

Restates kalman/enhanced_aircraft_kalman_tracker.py (AircraftKalmanTracker, :7-405) and
kalman/enhanced_multi_target_tracker.py (EnhancedMultiTargetTracker, :4-304) with the general dense
8x8 filter algebra of the reference (no structural shortcuts), including its quirks:

  * greedy (not Hungarian) association over candidates iou >= thr in descending IoU order
    (enhanced_multi_target_tracker.py:234-270); ties resolve to the lowest flat (det, trk) index here
    (np.argsort on exact ties is unspecified in the reference);
  * the extra predict() on the first lost frame via get_track_info -> get_lost_prediction ->
    enhanced_long_term_predict(1) (enhanced_aircraft_kalman_tracker.py:216-217, :319-333, :354);
  * the high-confidence long-term branch adds mean-velocity * k on top of an already advanced state
    (:226-236);
  * mixed precision: bbox_to_state and the detection-side area are evaluated in the dtype of the
    incoming detection scalars (float32 when they come from ``boxes.xyxy.cpu().numpy()``,
    kalman/aircraft_detection_tracking.py:101-106), everything else in float64.
print() side effects of the reference are omitted; the ``stats`` counters are kept.
"""
from __future__ import annotations

from collections import deque

import numpy as np

_F = np.eye(8)
_F[0, 4] = _F[1, 5] = _F[2, 6] = _F[3, 7] = 1.0                      # :50-54
_H = np.zeros((4, 8)); _H[0, 0] = _H[1, 1] = _H[2, 2] = _H[3, 3] = 1.0  # :57-61
_Q = np.diag([0.1, 0.1, 0.01, 0.01, 0.1, 0.1, 0.001, 0.001])         # :64-68
_R = np.eye(4) * 10.0                                                # :71
_P0 = np.diag([50.0, 50.0, 50.0, 50.0, 100.0, 100.0, 1.0, 1.0])      # :44-47


def _to_state(bbox):
    """bbox_to_state (:103-119) in the dtype of the incoming scalars."""
    b = np.asarray(bbox)
    if b.dtype not in (np.float32, np.float64):
        b = b.astype(np.float64)
    two = b.dtype.type(2.0)
    return np.array([(b[0] + b[2]) / two, (b[1] + b[3]) / two, b[2] - b[0], b[3] - b[1]], dtype=b.dtype)


def _to_bbox(x):
    """state_to_bbox (:121-135)."""
    return np.array([x[0] - x[2] / 2.0, x[1] - x[3] / 2.0, x[0] + x[2] / 2.0, x[1] + x[3] / 2.0])


class Track:
    def __init__(self, bbox, track_id, max_lost_frames):
        self.track_id = track_id
        self.age, self.hits, self.hit_streak, self.time_since_update = 0, 1, 1, 0
        self.x = np.zeros(8)
        self.P = _P0.copy()
        z = _to_state(bbox)
        self.x[:4] = z
        self.trajectory = deque(maxlen=150)
        self.velocities = deque(maxlen=50)
        self.velocity_avg = np.zeros(2)
        self.velocity_std = np.zeros(2)
        self.direction = 0.0
        self.speed = 0.0
        self.stability_score = 0.0
        self.prediction_confidence = 0.0
        self.is_lost, self.lost_frames = False, 0
        self.max_lost_frames = max_lost_frames
        self.trajectory.append((float(z[0]), float(z[1])))

    # ---- :137-182
    def analyze(self):
        if len(self.velocities) < 5:
            return
        v = np.array(self.velocities)
        self.velocity_avg = v.mean(0)
        self.velocity_std = v.std(0)
        ax, ay = self.velocity_avg
        self.speed = np.sqrt(ax ** 2 + ay ** 2)
        self.direction = np.arctan2(ay, ax)
        speed_stab = 1.0 / (1.0 + np.mean(self.velocity_std))
        d = np.arctan2(v[:, 1], v[:, 0])
        ch = np.diff(d)
        ch = np.array([c if abs(c) < np.pi else c - 2 * np.pi * np.sign(c) for c in ch])
        dir_cons = 1.0 / (1.0 + np.std(ch) * 10)
        self.stability_score = (speed_stab + dir_cons) / 2.0
        self.prediction_confidence = self.stability_score * min(len(self.velocities) / 30.0, 1.0)

    # ---- :184-203
    def predict(self):
        self.x = _F @ self.x
        self.P = _F @ self.P @ _F.T + _Q
        self.age += 1
        self.time_since_update += 1
        self.trajectory.append((self.x[0], self.x[1]))
        return _to_bbox(self.x)

    # ---- :205-247
    def long_term(self, k):
        if k <= 1:
            return self.predict(), 1.0
        self.analyze()
        if self.prediction_confidence > 0.3:
            s = self.x.copy()
            s[0] += self.velocity_avg[0] * k
            s[1] += self.velocity_avg[1] * k
            conf = self.prediction_confidence * max(0.1, 1.0 - k / self.max_lost_frames)
        else:
            s = self.x.copy()
            for _ in range(k):
                s = _F @ s
            conf = max(0.1, 1.0 - k / (self.max_lost_frames * 0.5))
        return _to_bbox(s), conf

    # ---- :249-297
    def update(self, bbox):
        self.time_since_update = 0
        self.hits += 1
        self.hit_streak += 1
        if self.is_lost:
            self.is_lost, self.lost_frames = False, 0
        z = _to_state(bbox).astype(np.float64)
        y = z - _H @ self.x
        S = _H @ self.P @ _H.T + _R
        K = self.P @ _H.T @ np.linalg.inv(S)
        self.x = self.x + K @ y
        self.P = (np.eye(8) - K @ _H) @ self.P
        self.velocities.append(self.x[4:6].copy())
        self.trajectory.append((self.x[0], self.x[1]))
        self.analyze()

    # ---- :299-317
    def mark_lost(self):
        if not self.is_lost:
            self.is_lost, self.lost_frames = True, 0
        self.lost_frames += 1
        self.hit_streak = 0

    # ---- :335-383
    def info(self):
        predicted = self.time_since_update > 0
        if predicted:
            if self.is_lost:
                bbox, conf = self.long_term(self.lost_frames)      # get_lost_prediction :319-333
            else:
                bbox = _to_bbox(self.x)
                conf = max(0.3, 1.0 - self.time_since_update / 60.0)
            status = "predicted"
        else:
            bbox, conf, status = _to_bbox(self.x), 1.0, "detected"
        return {
            "track_id": self.track_id, "bbox": bbox, "confidence": conf, "status": status,
            "age": self.age, "hits": self.hits, "hit_streak": self.hit_streak,
            "time_since_update": self.time_since_update, "lost_frames": self.time_since_update,
            "is_lost": predicted, "trajectory": list(self.trajectory)[-30:],
            "velocity": self.x[4:6], "motion_confidence": self.prediction_confidence,
            "is_stable_motion": self.stability_score > 0.5, "speed": self.speed, "direction": self.direction,
        }

    # ---- :385-405
    def should_delete(self, max_lost):
        if self.time_since_update > max_lost:
            return True
        if self.age < 5 and self.hit_streak == 0 and self.time_since_update > 15:
            return True
        if self.age < 10 and self.hit_streak <= 1 and self.time_since_update > 30:
            return True
        return False


def iou_matrix(dets, trk_boxes):
    """_calculate_iou_matrix / _calculate_iou (enhanced_multi_target_tracker.py:180-232), vectorised.
    dets: (D,>=4) in their own dtype; trk_boxes: (T,4) float64.  Returns (D,T) float64."""
    d = np.asarray(dets)[:, :4]
    if d.dtype not in (np.float32, np.float64):
        d = d.astype(np.float64)
    t = np.asarray(trk_boxes, np.float64).reshape(-1, 4)
    x1 = np.maximum(d[:, None, 0].astype(np.float64), t[None, :, 0])
    y1 = np.maximum(d[:, None, 1].astype(np.float64), t[None, :, 1])
    x2 = np.minimum(d[:, None, 2].astype(np.float64), t[None, :, 2])
    y2 = np.minimum(d[:, None, 3].astype(np.float64), t[None, :, 3])
    empty = (x2 <= x1) | (y2 <= y1)
    inter = (x2 - x1) * (y2 - y1)
    a1 = ((d[:, 2] - d[:, 0]) * (d[:, 3] - d[:, 1])).astype(np.float64)     # detection area in det dtype
    a2 = (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
    union = a1[:, None] + a2[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(empty | (union <= 0), 0.0, inter / union)
    return iou


def greedy_match(iou, thr):
    """_solve_assignment_problem (:234-270): candidates iou >= thr, descending IoU, greedy unique pairs.
    Ties -> lowest flat index (stable sort over np.where's row-major order)."""
    if iou.size == 0:
        return []
    di, ti = np.where(iou >= thr)
    if len(di) == 0:
        return []
    order = np.argsort(-iou[di, ti], kind="stable")
    used_d, used_t, out = set(), set(), []
    for k in order:
        d, t = int(di[k]), int(ti[k])
        if d not in used_d and t not in used_t:
            out.append((d, t)); used_d.add(d); used_t.add(t)
    return out


class MultiTracker:
    """EnhancedMultiTargetTracker (enhanced_multi_target_tracker.py:15-132, :288-304)."""

    def __init__(self, max_lost_frames=450, min_hits=3, iou_threshold=0.3):
        self.trackers = []
        self.max_lost_frames, self.min_hits, self.iou_threshold = max_lost_frames, min_hits, iou_threshold
        self.frame_count, self.next_track_id = 0, 1
        self.stats = {"total_tracks_created": 0, "total_tracks_terminated": 0, "current_active_tracks": 0,
                      "long_term_predictions": 0, "successful_recoveries": 0}
        self.last_min_iou_gap = np.inf   # diagnostic for near-tie fixtures: all candidates
        self.min_competing_gap = np.inf  # smallest IoU difference between two candidates that share a detection or a track
        self.min_thr_gap = np.inf        # smallest |IoU - threshold| over overlapping pairs

    def update(self, detections):
        self.frame_count += 1
        pred = [t.predict() for t in self.trackers]
        D, T = len(detections), len(self.trackers)
        if D > 0 and T > 0:
            iou = iou_matrix(detections, np.array(pred))
            cand = np.sort(iou[iou >= self.iou_threshold])
            if len(cand) > 1:
                self.last_min_iou_gap = min(self.last_min_iou_gap, float(np.diff(cand).min()))
            self._fixture_gaps(iou)
            matched = greedy_match(iou, self.iou_threshold)
            md, mt = {m[0] for m in matched}, {m[1] for m in matched}
            um_d = [d for d in range(D) if d not in md]
            um_t = [t for t in range(T) if t not in mt]
        else:
            matched, um_d, um_t = [], list(range(D)), list(range(T))
        for d, t in matched:
            trk = self.trackers[t]
            was_lost = trk.is_lost
            trk.update(np.asarray(detections[d])[:4])
            if was_lost:
                self.stats["successful_recoveries"] += 1
        for t in um_t:
            self.trackers[t].mark_lost()
        for d in um_d:
            self.trackers.append(Track(np.asarray(detections[d])[:4], f"T{self.next_track_id:03d}", self.max_lost_frames))
            self.next_track_id += 1
            self.stats["total_tracks_created"] += 1
        alive = []
        for trk in self.trackers:
            if trk.should_delete(self.max_lost_frames):
                self.stats["total_tracks_terminated"] += 1
            else:
                alive.append(trk)
        self.trackers = alive
        self.stats["current_active_tracks"] = len(alive)
        out = []
        for trk in self.trackers:
            if trk.hit_streak >= self.min_hits or self.frame_count <= self.min_hits or trk.is_lost:
                info = trk.info()
                out.append(info)
                if info["status"] == "predicted" and info["lost_frames"] > 30:
                    self.stats["long_term_predictions"] += 1
        return out

    def _fixture_gaps(self, iou):
        """How far the frame is from a decision an fp32 IoU could flip: only candidates that COMPETE (same detection or same
        track) can reorder the greedy walk; a pair near the threshold can enter or leave the candidate set."""
        pos = iou[iou > 0]
        if pos.size:
            self.min_thr_gap = min(self.min_thr_gap, float(np.abs(pos - self.iou_threshold).min()))
        m = np.where(iou >= self.iou_threshold, iou, np.nan)
        for axis in (0, 1):
            a = np.sort(m, axis=axis)
            d = np.diff(a, axis=axis)
            if np.isfinite(d).any():
                self.min_competing_gap = min(self.min_competing_gap, float(np.nanmin(d)))

    def get_statistics(self):
        return {**self.stats, "frame_count": self.frame_count,
                "tracker_details": [{"track_id": t.track_id, "age": t.age, "hits": t.hits,
                                     "lost_frames": t.lost_frames, "is_lost": t.is_lost,
                                     "confidence": t.prediction_confidence} for t in self.trackers]}
