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

    def _should_global_reset(self):                       # :119-146
        info = self.frame_motion_info
        if not info or not info["should_reset"]:
            return False
        if len(self.detection_stability_history) >= 5:
            recent = list(self.detection_stability_history)[-5:]
            if np.std(recent) / (np.mean(recent) + 1) > 0.5:
                return True
        if len(self.global_motion_history) >= 3 and np.mean(list(self.global_motion_history)[-3:]) > 30.0:
            return True
        return info["magnitude"] > 60.0

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

    def update(self, detections, frame=None):             # :75-117 + :168-238; frame drives the global detector (:92-114)
        from collections import deque

        self.frame_count += 1
        self.stats["total_frames"] += 1
        if not hasattr(self, "motion_detector"):
            self.motion_detector = GlobalMotionDetector()
            self.global_motion_history, self.detection_stability_history = deque(maxlen=20), deque(maxlen=10)
            self.stats.setdefault("global_motion_events", 0); self.stats.setdefault("global_resets", 0)
            self.frame_motion_info = None
        global_motion = False
        if frame is not None:
            is_motion, mag, vec, should_reset = self.motion_detector.detect_motion(frame)
            self.frame_motion_info = {"is_motion": is_motion, "magnitude": mag, "should_reset": should_reset}
            self.global_motion_history.append(mag)
            if should_reset:
                global_motion = True
                self.stats["global_motion_events"] += 1
        self.detection_stability_history.append(len(detections))
        if global_motion and self._should_global_reset():  # :116-117, :148-166: every track dropped, one new track per detection
            self.stats["global_resets"] += 1
            self.trackers = [self._new(d[:4]) for d in detections]
            return [t.info() for t in self.trackers]
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


class GlobalMotionDetector:
    """camera_motion_compensation/global_motion_detector.py, method 'optical_flow' (the tracker's default, :30-31 of the multi
    tracker): corners of the PREVIOUS frame tracked into the current one, the global vector is the mean flow of the 75 % of
    points closest to the median flow; motion above 30 px, reset above 50 px or above 45 px when the last three vectors point the
    same way (:113-184, consistency :262-280).  The image operations are OpenCV's (third party): cvtColor, goodFeaturesToTrack,
    calcOpticalFlowPyrLK, called with the reference's parameters."""

    def __init__(self):
        from collections import deque

        self.prev_gray = None
        self.motion_vectors = deque(maxlen=5)
        self.motion_threshold, self.reset_threshold, self.consistency_threshold = 30.0, 50.0, 0.7
        self.stats = {"total_detections": 0, "motion_events": 0, "reset_triggers": 0}

    def detect_motion(self, frame):
        import cv2

        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        if self.prev_gray is None:
            self.prev_gray = gray
            return False, 0.0, np.array([0.0, 0.0]), False
        res = self._flow(gray)
        self.prev_gray = gray.copy()
        self.stats["total_detections"] += 1
        self.stats["motion_events"] += int(res[0]); self.stats["reset_triggers"] += int(res[3])
        return res

    def _flow(self, gray):
        import cv2

        none = (False, 0.0, np.array([0.0, 0.0]), False)
        corners = cv2.goodFeaturesToTrack(self.prev_gray, maxCorners=200, qualityLevel=0.01, minDistance=15, blockSize=7)
        if corners is None or len(corners) < 20:
            return none
        nxt, status, _ = cv2.calcOpticalFlowPyrLK(self.prev_gray, gray, corners, None, winSize=(21, 21), maxLevel=3,
                                                  criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01))
        if status is None:
            return none
        good = status.flatten() == 1
        if np.sum(good) < 10:
            return none
        mv = nxt[good].reshape(-1, 2) - corners[good].reshape(-1, 2)
        if len(mv) > 8:
            dist = np.linalg.norm(mv - np.median(mv, axis=0), axis=1)
            inl = dist < np.percentile(dist, 75)
            if np.sum(inl) > 5:
                g = np.mean(mv[inl], axis=0)
                mag = np.linalg.norm(g)
                self.motion_vectors.append(g)
                is_motion, should_reset = mag > self.motion_threshold, mag > self.reset_threshold
                if len(self.motion_vectors) >= 3:
                    ang = [np.arctan2(v[1], v[0]) for v in list(self.motion_vectors)[-3:]]
                    diffs = []
                    for i in range(1, len(ang)):
                        d = abs(ang[i] - ang[i - 1])
                        diffs.append(2 * np.pi - d if d > np.pi else d)
                    if max(0.0, 1.0 - np.mean(diffs) / np.pi) > self.consistency_threshold and is_motion:
                        should_reset = should_reset or mag > self.motion_threshold * 1.5
                return is_motion, mag, g, should_reset
        return none
