"""Oracle (numpy float64, CPU) for camera_motion_compensation/motion_reset_kalman_tracker.py -- test infrastructure only.

SURVEY.md 8f N1 (the first "next" row): MotionResetKalmanTracker (:16-355) = the hot path's AircraftKalmanTracker
(oracle.tracker.Track) plus per-track reset heuristics.  Restated with the reference's quirks:

  * the three detectors look at the LAST THREE stored centres (+ the new one), the previous bbox, and run only outside the
    15-frame cooldown counted in ``age`` (:200-203); _detect_position_jump appends a motion score as a side effect (:88-89);
  * reset confidence = mean of the triggered factors, x1.5 when motion consistency < 0.3, x0.8 when a reset happened less than
    50 frames ago; reset iff confidence > 1.0 (:223-241);
  * a reset overwrites x[:4], zeroes the velocity, scales P[4:,4:] by 100 and P[:4,:4] by 5 (the cross blocks are left as
    they are), clears the histories, bumps hits / hit_streak, zeroes time_since_update -- and does NOT clear is_lost /
    lost_frames as a normal update does (:246-279);
  * ``position_history`` is shared with the base class by accident of naming: every normal update appends the filtered
    centre (base :289-290) and then the raw detection centre (:284), so "the last three positions" alternate between the two;
  * predict() blends the predicted centre with the last stored centre for 10 frames after a reset (:300-321); the blended box
    is only what predict() returns -- the state is not touched.
Pinned by tests/golden/motion_reset.npz (tests/golden/make_golden.py motion_reset, the unmodified reference).
"""
from __future__ import annotations

from collections import deque

import numpy as np

from .tracker import Track, _to_state


def _center(b):
    return np.array([(b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0])


def _size(b):
    return np.array([b[2] - b[0], b[3] - b[1]])


class MotionResetTrack(Track):
    def __init__(self, bbox, track_id, max_lost_frames=150):
        super().__init__(bbox, track_id, max_lost_frames)
        self.position_history = deque(maxlen=8)             # :41
        self.bbox_history = deque(maxlen=5)                 # :43
        self.jump_threshold, self.velocity_threshold, self.size_change_threshold, self.reset_cooldown = 40.0, 60.0, 0.3, 15   # :46-49
        self.reset_count, self.last_reset_frame = 0, -999   # :52-53
        self.reset_log = []
        self.motion_scores = deque(maxlen=10)               # :55
        self.motion_consistency = 0.0
        self.position_history.append(_center(bbox))
        self.bbox_history.append(list(bbox))

    # ---- :78-94
    def _position_jump(self, c):
        if len(self.position_history) < 2:
            return False, 0.0
        avg = np.mean(list(self.position_history)[-3:], axis=0)
        dist = float(np.linalg.norm(c - avg))
        self.motion_scores.append(min(dist / self.jump_threshold, 3.0))
        return dist > self.jump_threshold, dist

    # ---- :96-121
    def _velocity_change(self, c):
        if len(self.position_history) < 3:
            return False, 0.0
        pos = list(self.position_history)[-3:] + [c]
        v = [float(np.linalg.norm(pos[i] - pos[i - 1])) for i in range(1, len(pos))]
        change = abs(v[-1] - float(np.mean(v[:-1])))
        return change > self.velocity_threshold, change

    # ---- :123-142
    def _size_change(self, bbox):
        if len(self.bbox_history) < 2:
            return False, 0.0
        prev = np.maximum(_size(self.bbox_history[-1]), 1.0)
        ratio = _size(bbox) / prev
        m = float(max(abs(ratio[0] - 1.0), abs(ratio[1] - 1.0)))
        return m > self.size_change_threshold, m

    # ---- :144-159
    def _consistency(self):
        if len(self.motion_scores) < 3:
            return 0.0
        s = list(self.motion_scores)
        mean = float(np.mean(s))
        return max(0.0, 1.0 - float(np.var(s)) / (mean + 0.1)) if mean > 0 else 1.0

    # ---- :161-244
    def should_reset(self, bbox):
        since = self.age - self.last_reset_frame
        if since < self.reset_cooldown:
            return False, 0.0
        c = _center(bbox)
        f = []
        j, d = self._position_jump(c)
        if j:
            f.append(min(d / self.jump_threshold, 2.0))
        v, dv = self._velocity_change(c)
        if v:
            f.append(min(dv / self.velocity_threshold, 2.0))
        s, ds = self._size_change(bbox)
        if s:
            f.append(ds / self.size_change_threshold)
        if not f:
            return False, 0.0
        conf = float(np.mean(f))
        self.motion_consistency = self._consistency()
        if self.motion_consistency < 0.3:
            conf *= 1.5
        if self.reset_count > 0 and since < 50:
            conf *= 0.8
        return conf > 1.0, conf

    # ---- :246-279
    def reset(self, bbox, conf):
        self.reset_count += 1
        self.last_reset_frame = self.age
        self.reset_log.append((self.age, conf, self.motion_consistency))
        self.x[:4] = _to_state(bbox)
        self.x[4:] = 0
        self.P[4:, 4:] *= 100.0
        self.P[:4, :4] *= 5.0
        c = _center(bbox)
        self.trajectory.clear()
        self.trajectory.append((c[0], c[1]))
        self.velocities.clear()
        self.position_history.clear()
        self.position_history.append(c)
        self.motion_scores.clear()
        self.hits += 1
        self.hit_streak += 1
        self.time_since_update = 0

    # ---- :281-298
    def update(self, bbox):
        do, conf = self.should_reset(bbox)
        if do:
            self.reset(bbox, conf)
        else:
            super().update(bbox)
            # the base class keeps a ``position_history`` of its own (enhanced_aircraft_kalman_tracker.py:80, :289-290) which the
            # subclass shadows with its 8-deep deque: a normal update therefore stores the FILTERED centre here, and the
            # raw detection centre right after it
            self.position_history.append(self.x[:2].copy())
        self.position_history.append(_center(bbox))
        self.bbox_history.append(list(bbox))

    # ---- :300-321
    def predict(self):
        pb = super().predict()
        since = self.age - self.last_reset_frame
        if since < 10 and len(self.position_history) > 0:
            last = self.position_history[-1]
            blend = min(since / 10.0, 1.0)
            c = (1 - blend) * last + blend * _center(pb)
            sz = _size(pb)
            pb = np.array([c[0] - sz[0] / 2, c[1] - sz[1] / 2, c[0] + sz[0] / 2, c[1] + sz[1] / 2])
        return pb

    # ---- :323-342 (numeric fields only)
    def info(self):
        d = super().info()
        d["reset_count"] = self.reset_count
        d["frames_since_reset"] = self.age - self.last_reset_frame
        d["motion_consistency"] = self.motion_consistency
        return d


# ------------------------------------------------------------------------------------------------
# camera_motion_compensation/motion_compensated_multi_tracker.py (:18-394), the path taken when no frame is passed
# (global motion detection needs the optical-flow GlobalMotionDetector, global_motion_detector.py: not restated).
# Differences from the hot path's EnhancedMultiTargetTracker that matter for parity:
#   * its own association (:240-283): candidates iou > thr (strict), sorted as (iou, d, t) tuples in DESCENDING order, so
#     exact ties go to the larger detection index, then the larger tracker index;
#   * tracks are MotionResetKalmanTrackers created with track_id=None (random uuid ids in the reference: compare by list
#     position); every live track is reported (no min_hits gate);
#   * stats: individual_resets (+1 per matched update that reset), tracking_recoveries (+1 per deleted track that had reset).
# ------------------------------------------------------------------------------------------------
def _iou(a, b):
    x1, y1, x2, y2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    if x2 <= x1 or y2 <= y1:
        return 0.0
    inter = (x2 - x1) * (y2 - y1)
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return 0.0 if union <= 0 else inter / union


class MotionCompensatedMultiTracker:
    def __init__(self, max_lost_frames=150, min_hits=1, iou_threshold=0.1):
        self.max_lost_frames, self.min_hits, self.iou_threshold = max_lost_frames, min_hits, iou_threshold
        self.trackers = []
        self.frame_count = 0
        self.stats = {"total_frames": 0, "individual_resets": 0, "tracking_recoveries": 0}
        self._next = 1

    def _new(self, bbox):
        t = MotionResetTrack(bbox, f"N{self._next:04d}", self.max_lost_frames)
        self._next += 1
        return t

    def associate(self, dets, preds):                     # :240-283
        if len(dets) == 0:
            return [], [], list(range(len(preds)))
        if len(preds) == 0:
            return [], list(range(len(dets))), []
        cand = []
        for d, det in enumerate(dets):
            for t, p in enumerate(preds):
                v = _iou(det[:4], p)
                if v > self.iou_threshold:
                    cand.append((v, d, t))
        cand.sort(reverse=True)
        used_d, used_t, matched = set(), set(), []
        for _, d, t in cand:
            if d not in used_d and t not in used_t:
                matched.append([d, t]); used_d.add(d); used_t.add(t)
        return matched, [d for d in range(len(dets)) if d not in used_d], [t for t in range(len(preds)) if t not in used_t]

    def update(self, detections):                         # :76-117 (frame is None) + :168-238
        self.frame_count += 1
        self.stats["total_frames"] += 1
        preds = [t.predict() for t in self.trackers]
        if len(detections) > 0 and len(self.trackers) > 0:
            matched, um_d, um_t = self.associate(detections, preds)
        else:
            matched, um_d, um_t = [], list(range(len(detections))), list(range(len(self.trackers)))
        for d, t in matched:
            before = self.trackers[t].reset_count
            self.trackers[t].update(detections[d][:4])
            self.stats["individual_resets"] += int(self.trackers[t].reset_count > before)
        for t in um_t:
            self.trackers[t].mark_lost()
        for d in um_d:
            self.trackers.append(self._new(detections[d][:4]))
        keep = []
        for t in self.trackers:
            if t.should_delete(self.max_lost_frames):
                self.stats["tracking_recoveries"] += int(t.reset_count > 0)
            else:
                keep.append(t)
        self.trackers = keep
        return [t.info() for t in self.trackers]
