"""Host-side mirror of the project's tracker API over the CUDA Kalman track bank (csrc/tracker.cu).

  EnhancedMultiTargetTracker   same constructor, ``update(detections) -> list[dict]``, ``stats``,
                               ``frame_count``, ``next_track_id``, ``trackers``, ``get_statistics()`` as
                               kalman/enhanced_multi_target_tracker.py:4-304 (one video stream).
  TrackerBank                  the batched form the B200 pipeline uses: S independent streams advanced by
                               one launch sequence per frame, detections and results resident on the GPU.

All Kalman / association / lifecycle arithmetic runs in the CUDA library; this module only marshals
arguments and formats the output dicts (get_track_info, enhanced_aircraft_kalman_tracker.py:335-383).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib

TRACK_COLS, TRAJ_LEN = _lib.TRACK_COLS, _lib.TRAJ_LEN
_STAT_KEYS = ("total_tracks_created", "total_tracks_terminated", "current_active_tracks",
              "long_term_predictions", "successful_recoveries")


class TrackerBank:
    """S independent multi-target trackers (one per video stream) resident on one GPU.

    ``capacity`` bounds the simultaneously live tracks per stream (the reference's list is unbounded,
    enhanced_multi_target_tracker.py:92-101): ``grow()`` enlarges the bank in place, ``stats_async()`` /
    ``export()`` expose the ``dropped`` counter, and the callers in this package (EnhancedMultiTargetTracker,
    DetectTrackPipeline) grow ahead of need and raise if a detection was ever dropped.
    ``max_out`` (default: capacity) is the number of rows per stream the output block holds."""

    def __init__(self, n_streams, capacity=256, max_dets=300, max_lost_frames=450, min_hits=3, iou_threshold=0.3, device=None,
                 max_out=None, mode=0):
        """mode 1: the rules of camera_motion_compensation's MotionCompensatedMultiTracker / MotionResetKalmanTracker."""
        self.device = device or _lib.require_cuda()
        self.lib = _lib.load()
        self.mode = int(mode)
        self.S, self.capacity, self.max_dets = int(n_streams), int(capacity), int(max_dets)
        self.max_lost_frames, self.min_hits, self.iou_threshold = int(max_lost_frames), int(min_hits), float(iou_threshold)
        self._follow_capacity = max_out is None
        self.max_out = self.capacity if max_out is None else int(max_out)
        self._h = C.c_void_p()
        _lib.check(self.lib.b2_tracker_create_ex(self.S, self.capacity, self.max_dets, self.max_lost_frames, self.min_hits,
                                                 self.iou_threshold, self.mode, C.byref(self._h)))
        self._alloc_outputs()

    def _alloc_outputs(self):
        import torch

        self.rows = torch.zeros((self.S, self.max_out, TRACK_COLS), dtype=torch.float32, device=self.device)
        self.counts = torch.zeros((self.S,), dtype=torch.int32, device=self.device)
        self.traj = self.traj_len = None
        self.extra = torch.zeros((self.S, self.max_out, 4), dtype=torch.float32, device=self.device) if self.mode else None
        self.stats_dev = torch.zeros((self.S, 8), dtype=torch.int64, device=self.device)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b2_tracker_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        _lib.check(self.lib.b2_tracker_reset(self._h, _lib.stream_ptr()))

    def grow(self, new_capacity):
        """Enlarge every stream's bank to ``new_capacity`` slots (state, ids and slot indices are kept; synchronises)."""
        _lib.check(self.lib.b2_tracker_grow(self._h, int(new_capacity), _lib.stream_ptr()))
        self.capacity = int(new_capacity)
        if self._follow_capacity:
            self.max_out = self.capacity
            self._alloc_outputs()

    def update(self, dets, det_counts, with_trajectory=True, stream=None):
        """dets: CUDA float32 [S][max_dets][cols>=4] rows x1,y1,x2,y2,...; det_counts: CUDA int32 [S].
        Returns (rows [S][max_out][20], counts [S]) device tensors (views of internal buffers)."""
        import torch

        assert dets.is_cuda and dets.is_contiguous() and dets.shape[0] == self.S and dets.shape[1] == self.max_dets
        if with_trajectory and self.traj is None:
            self.traj = torch.zeros((self.S, self.max_out, TRAJ_LEN, 2), dtype=torch.float32, device=self.device)
            self.traj_len = torch.zeros((self.S, self.max_out), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.b2_tracker_update_ex(self._h, _lib.ptr(dets), dets.shape[2], _lib.ptr(det_counts), _lib.ptr(self.rows),
                                                 _lib.ptr(self.counts), _lib.ptr(self.traj) if with_trajectory else None,
                                                 _lib.ptr(self.traj_len) if with_trajectory else None, _lib.ptr(self.extra), self.max_out,
                                                 _lib.stream_ptr(stream)))
        return self.rows, self.counts

    def stats_async(self, stream=None):
        """[S][8] int64 device snapshot {created, terminated, active, long_term, recoveries, dropped, 0, 0}, stream-ordered."""
        _lib.check(self.lib.b2_tracker_stats(self._h, _lib.ptr(self.stats_dev), _lib.stream_ptr(stream)))
        return self.stats_dev

    def predict_only(self, stream=None):
        """AircraftKalmanTracker.predict alone on every live track."""
        _lib.check(self.lib.b2_tracker_bank_predict(self._h, _lib.stream_ptr(stream)))

    def export(self, stream_idx=0):
        """Dense state of one stream (synchronises): x (n,8), P (n,8,8), meta (n,8) int32, stats (8,) int64."""
        cap = self.capacity
        x = np.zeros((cap, 8), np.float32)
        P = np.zeros((cap, 64), np.float32)
        meta = np.zeros((cap, 8), np.int32)
        n = C.c_int32()
        stats = (C.c_longlong * 8)()
        _lib.check(self.lib.b2_tracker_export(self._h, stream_idx, x.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p),
                                              meta.ctypes.data_as(C.c_void_p), C.byref(n), stats))
        k = n.value
        order = np.argsort(meta[:k, 0], kind="stable")          # reference list order == ascending track id
        return x[:k][order], P[:k].reshape(k, 8, 8)[order], meta[:k][order], np.array(list(stats), np.int64)

    def export_motion(self, stream_idx=0):
        """prediction_confidence of one stream's live tracks, ascending track id (synchronises)."""
        m = np.zeros((self.capacity, 8), np.float32)
        n = C.c_int32()
        _lib.check(self.lib.b2_tracker_export_motion(self._h, stream_idx, m.ctypes.data_as(C.c_void_p), C.byref(n)))
        m = m[:n.value]
        return m[np.argsort(m[:, 0].copy().view(np.int32), kind="stable"), 6]

    def export_reset(self, stream_idx=0):
        """mode 1: (n, 7) {id, reset_count, last_reset_frame, motion_consistency, len(position_history), len(motion_scores),
        len(bbox_history)} of one stream's live tracks, ascending track id (synchronises)."""
        m = np.zeros((self.capacity, 8), np.float32)
        n = C.c_int32()
        _lib.check(self.lib.b2_tracker_export_reset(self._h, stream_idx, m.ctypes.data_as(C.c_void_p), C.byref(n)))
        m = m[:n.value]
        iv = m.view(np.int32)
        order = np.argsort(iv[:, 0], kind="stable")
        out = np.stack([iv[:, 0], iv[:, 1], iv[:, 2], m[:, 3], iv[:, 4], iv[:, 5], iv[:, 6]], 1).astype(np.float64)
        return out[order]

    @staticmethod
    def bytes_per_track():
        a, b = C.c_int(), C.c_int()
        _lib.load().b2_tracker_bytes_per_track(C.byref(a), C.byref(b))
        return a.value, b.value


def rows_to_dicts(rows, traj=None, traj_len=None):
    """(n, 20) float32 rows (+ trajectories) of one stream -> list of get_track_info dicts, ascending id."""
    rows = np.ascontiguousarray(rows, np.float32)
    irows = rows.view(np.int32)
    order = np.argsort(irows[:, 0], kind="stable")
    out = []
    for k in order:
        r, ir = rows[k], irows[k]
        predicted = bool(ir[6])
        d = {
            "track_id": f"T{int(ir[0]):03d}",
            "bbox": r[1:5].astype(np.float64),
            "confidence": float(r[5]),
            "status": "predicted" if predicted else "detected",
            "age": int(ir[7]), "hits": int(ir[8]), "hit_streak": int(ir[9]),
            "time_since_update": int(ir[10]), "lost_frames": int(ir[11]), "is_lost": bool(ir[12]),
            "trajectory": [] if traj is None else [(float(a), float(b)) for a, b in traj[k][:int(traj_len[k])]],
            "velocity": r[13:15].astype(np.float64),
            "motion_confidence": float(r[15]), "is_stable_motion": bool(ir[16]),
            "speed": float(r[17]), "direction": float(r[18]),
        }
        out.append(d)
    return out


class _TrackView:
    """Read-only view of one track, with the attributes callers of ``tracker.trackers`` read
    (kalman/enhanced_multi_target_tracker.py:288-304, camera_motion_compensation subclasses)."""

    def __init__(self, x, P, meta):
        self.track_id = f"T{int(meta[0]):03d}"
        self.x, self.P = x.astype(np.float64), P.astype(np.float64)
        self.age, self.hits, self.hit_streak = int(meta[1]), int(meta[2]), int(meta[3])
        self.time_since_update, self.lost_frames, self.is_lost = int(meta[4]), int(meta[5]), bool(meta[6])


class EnhancedMultiTargetTracker:
    """Drop-in for kalman/enhanced_multi_target_tracker.py:4 (single stream).

    The reference's track list is unbounded; the bank behind this class starts at ``capacity`` slots and doubles
    (``TrackerBank.grow``) before a frame could overflow it, so no detection is ever dropped.  ``max_dets`` bounds the
    detections of one frame (the NMS cap of the detector, 300); more raise ``ValueError``.
    """

    def __init__(self, max_lost_frames=450, min_hits=3, iou_threshold=0.3, capacity=512, max_dets=300, verbose=False):
        import torch

        self.max_lost_frames, self.min_hits, self.iou_threshold = max_lost_frames, min_hits, iou_threshold
        self.bank = TrackerBank(1, capacity, max_dets, max_lost_frames, min_hits, iou_threshold)
        self.frame_count = 0
        self.next_track_id = 1
        self.stats = {k: 0 for k in _STAT_KEYS}
        self.verbose = verbose
        self._dets = torch.zeros((1, max_dets, 4), dtype=torch.float32, device=self.bank.device)
        self._host = torch.zeros((max_dets, 4), dtype=torch.float32).pin_memory()
        self._cnt = torch.zeros((1,), dtype=torch.int32, device=self.bank.device)

    def update(self, detections):
        """detections: list of [x1, y1, x2, y2, conf] -> list of track-info dicts (ascending track id)."""
        import torch

        n = len(detections)
        if n > self.bank.max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self.bank.max_dets}")
        # every detection may found a track: make room first (the reference appends without bound, :92-101)
        need = self.stats["current_active_tracks"] + n
        if need > self.bank.capacity:
            self.bank.grow(min(65535, max(2 * self.bank.capacity, need)))
        if n:
            self._host[:n] = torch.as_tensor(np.asarray([list(d)[:4] for d in detections], dtype=np.float32))
            self._dets[0, :n].copy_(self._host[:n], non_blocking=True)
        self._cnt.fill_(n)
        rows, counts = self.bank.update(self._dets, self._cnt)
        k = int(counts[0].item())
        r = rows[0, :k].cpu().numpy()
        tr = self.bank.traj[0, :k].cpu().numpy()
        tl = self.bank.traj_len[0, :k].cpu().numpy()
        self._refresh_stats()
        return rows_to_dicts(r, tr, tl)

    def _refresh_stats(self):
        st = (C.c_longlong * 8)()
        _lib.check(self.bank.lib.b2_tracker_export(self.bank._h, 0, None, None, None, None, st))
        for k, v in zip(_STAT_KEYS, list(st)[:5]):
            self.stats[k] = int(v)
        self.frame_count, self.next_track_id = int(st[5]), int(st[6])
        if st[7]:      # cannot happen below 65535 live tracks: update() grows the bank first
            raise RuntimeError(f"track bank overflow: {int(st[7])} detections found no free slot (capacity={self.bank.capacity})")

    @property
    def trackers(self):
        x, P, meta, _ = self.bank.export(0)
        return [_TrackView(x[i], P[i], meta[i]) for i in range(len(x))]

    def get_statistics(self):
        """enhanced_multi_target_tracker.py:288-304."""
        x, P, meta, stats = self.bank.export(0)
        conf = self.bank.export_motion(0)
        d = dict(self.stats)
        d["frame_count"] = self.frame_count
        d["tracker_details"] = [{"track_id": f"T{int(m[0]):03d}", "age": int(m[1]), "hits": int(m[2]), "lost_frames": int(m[5]),
                                 "is_lost": bool(m[6]), "confidence": float(c)} for m, c in zip(meta, conf)]
        return d


class MotionCompensatedMultiTracker(EnhancedMultiTargetTracker):
    """Drop-in for camera_motion_compensation/motion_compensated_multi_tracker.py:18 on the CUDA track bank (mode 1):
    MotionResetKalmanTracker tracks (per-track jump / velocity / size reset detectors, cooldown, covariance rescale, blended
    association box), the subclass's own association (IoU > threshold, ties to the larger indices) and reporting of every
    live track.  ``update(detections, frame=None)``: the global camera-motion detector (optical flow on ``frame``,
    global_motion_detector.py) is not part of this path -- passing a frame raises."""

    def __init__(self, max_lost_frames=150, min_hits=1, iou_threshold=0.1, capacity=512, max_dets=300):
        import torch

        self.max_lost_frames, self.min_hits, self.iou_threshold = max_lost_frames, min_hits, iou_threshold
        self.bank = TrackerBank(1, capacity, max_dets, max_lost_frames, min_hits, iou_threshold, mode=1)
        self.frame_count = 0
        self.next_track_id = 1
        self.stats = {"total_frames": 0, "individual_resets": 0, "tracking_recoveries": 0, "global_resets": 0}
        self._active = 0
        self._dets = torch.zeros((1, max_dets, 4), dtype=torch.float32, device=self.bank.device)
        self._host = torch.zeros((max_dets, 4), dtype=torch.float32).pin_memory()
        self._cnt = torch.zeros((1,), dtype=torch.int32, device=self.bank.device)

    def update(self, detections, frame=None):
        import torch

        if frame is not None:
            raise NotImplementedError("global camera-motion detection (GlobalMotionDetector, optical flow) is not implemented; call update(detections)")
        n = len(detections)
        if n > self.bank.max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self.bank.max_dets}")
        if self._active + n > self.bank.capacity:
            self.bank.grow(min(65535, max(2 * self.bank.capacity, self._active + n)))
        if n:
            self._host[:n] = torch.as_tensor(np.asarray([list(d)[:4] for d in detections], dtype=np.float32))
            self._dets[0, :n].copy_(self._host[:n], non_blocking=True)
        self._cnt.fill_(n)
        rows, counts = self.bank.update(self._dets, self._cnt)
        k = int(counts[0].item())
        r = rows[0, :k].cpu().numpy()
        ex = self.bank.extra[0, :k].cpu().numpy()
        tr, tl = self.bank.traj[0, :k].cpu().numpy(), self.bank.traj_len[0, :k].cpu().numpy()
        st = (C.c_longlong * 8)()
        _lib.check(self.bank.lib.b2_tracker_export(self.bank._h, 0, None, None, None, None, st))
        self._active = int(st[2])
        self.frame_count, self.next_track_id = int(st[5]), int(st[6])
        self.stats.update(total_frames=self.frame_count, individual_resets=int(st[3]), tracking_recoveries=int(st[4]))
        if st[7]:
            raise RuntimeError(f"track bank overflow: {int(st[7])} detections found no free slot")
        out = rows_to_dicts(r, tr, tl)
        order = np.argsort(r.view(np.int32)[:, 0], kind="stable")
        for d, j in zip(out, order):
            d["reset_count"] = int(ex[j].view(np.int32)[0])
            d["frames_since_reset"] = int(ex[j].view(np.int32)[1])
            d["motion_consistency"] = float(ex[j][2])
        return out

    def get_statistics(self):
        d = dict(self.stats)
        d["frame_count"] = self.frame_count
        d["active_trackers"] = self._active
        return d


def direction_wrap(c):
    """The wrap rule of _calculate_direction_consistency (enhanced_aircraft_kalman_tracker.py:165-182)."""
    return c if abs(c) < math.pi else c - 2 * math.pi * (1 if c > 0 else -1 if c < 0 else 0)
