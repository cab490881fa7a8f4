"""Track overlay for the detect+track loop: host-side mirror of the reference's
`kalman/trajectory_visualizer.py:5-234` (`TrajectoryVisualizer.draw_tracks(image, tracks, detections, frame_info)`), consuming the
dicts `tracker.EnhancedMultiTargetTracker.update` returns (bbox, track_id, status, time_since_update, confidence, trajectory,
velocity).  Drawing is OpenCV on the host, as in the reference: one overlay per displayed frame, off the GPU hot path.

Built as a display list: `compile_frame` turns a frame's tracks into a flat list of primitives (outline / filled box, 30 % tint,
text, poly-line segment, arrow) in the reference's painting order, `rasterize` executes it.  The list is plain data, so a test can
compare it (and the pixels it produces) with the reference without looking at images by eye."""
from __future__ import annotations

import math
from typing import Iterable, NamedTuple, Optional, Sequence

import cv2
import numpy as np

# BGR, the reference's palette (trajectory_visualizer.py:12-20)
PALETTE = {
    "detected": (0, 255, 0),
    "predicted": (0, 165, 255),
    "lost": (0, 100, 255),
    "trajectory": (255, 255, 0),
    "velocity": (255, 0, 255),
    "text": (255, 255, 255),
    "background": (0, 0, 0),
}
_FLASH = (0, 220, 255)          # bright phase of a coasting track's blinking outline (:69-74)
_FONT = cv2.FONT_HERSHEY_SIMPLEX


class Prim(NamedTuple):
    """One drawing primitive.  kind: 'box' (thickness -1 = filled), 'tint' (30 % blend of a filled box), 'text', 'seg', 'arrow'."""
    kind: str
    a: tuple                    # first point / text origin
    b: tuple = ()               # second point (box, tint, seg, arrow)
    color: tuple = (0, 0, 0)
    thickness: int = 1
    text: str = ""
    scale: float = 0.0


def _text_extent(text: str, scale: float, thickness: int) -> tuple:
    return cv2.getTextSize(text, _FONT, scale, thickness)[0]


class TrajectoryVisualizer:
    """Same attributes and call as the reference class; `colors` overrides the palette."""

    def __init__(self, colors: Optional[dict] = None):
        self.colors = colors or dict(PALETTE)
        self.trajectory_length = 20
        self.velocity_scale = 5.0
        self.font = _FONT
        self.font_scale = 0.4
        self.font_thickness = 1
        self.frame_counter = 0

    # ---- display list -------------------------------------------------------------------------------------------
    def _tag(self, out: list, text: str, org: tuple, fill: tuple, ink: tuple, scale: float, thickness: int) -> None:
        """Text on a filled plate that hugs it by 2 px (:119-135, :152-157)."""
        tw, th = _text_extent(text, scale, thickness)
        x, y = org
        out.append(Prim("box", (x - 2, y - th - 2), (x + tw + 2, y + 2), fill, -1))
        out.append(Prim("text", (x, y), (), ink, thickness, text, scale))

    def _status_origin(self, text: str, x2: int, y1: int, shape: Sequence[int]) -> tuple:
        """Where the state caption goes: right of the box, flipped to its left / above when it would leave the image (:137-150)."""
        tw, _ = _text_extent(text, 0.35, 1)
        x, y = x2 + 20, y1 + 15
        if x + tw > shape[1]:
            x = x2 - tw - 20
        if y > shape[0]:
            y = y1 - 10
        return x, y

    def _compile_track(self, out: list, track: dict, shape: Sequence[int]) -> None:
        box = track["bbox"]
        tid = str(track["track_id"])
        coasting = track.get("status", "detected") == "predicted"
        since = int(track.get("time_since_update", 0))
        conf = float(track.get("confidence", 1.0))
        x1, y1, x2, y2 = (int(float(v)) for v in box[:4])
        if coasting:
            bright = (self.frame_counter // 6) % 2 == 0                 # blinks every six frames (:69-74)
            col = _FLASH if bright else self.colors["predicted"]
            out.append(Prim("box", (x1, y1), (x2, y2), col, 2 if bright else 1))
            out.append(Prim("tint", (x1, y1), (x2, y2), col))
            self._tag(out, f"ID:{tid} PRED({since})", (x2 + 15, y1 - 5), col, self.colors["text"], self.font_scale, self.font_thickness)
            caption = "⚠️ AI PREDICTION"
        else:
            col = self.colors["detected"]
            out.append(Prim("box", (x1, y1), (x2, y2), col, 1))
            self._tag(out, f"ID:{tid} TRACKING", (x2 + 15, y1 - 5), col, self.colors["text"], self.font_scale, self.font_thickness)
            caption = "✅ DETECTED"
        self._tag(out, caption, self._status_origin(caption, x2, y1, shape), col, (255, 255, 255), 0.35, 1)    # white whatever the palette says (:156)
        out.append(Prim("text", (x2 + 10, y2 + 10), (), self.colors["text"], 1, f"Conf: {conf:.2f}", 0.3))
        # trail: the last `trajectory_length` centres, older segments thinner (:160-172)
        trail = track.get("trajectory", [])
        if len(trail) >= 2:
            pts = np.array(trail[-self.trajectory_length:], dtype=np.int32)
            n = len(pts)
            for i in range(1, n):
                out.append(Prim("seg", tuple(int(v) for v in pts[i - 1]), tuple(int(v) for v in pts[i]), self.colors["trajectory"],
                                max(1, int(3 * (i / n)))))
        vx, vy = track.get("velocity", (0, 0))
        if math.sqrt(vx ** 2 + vy ** 2) > 1.0:                           # velocity arrow from the box centre (:174-184)
            cx, cy = int((box[0] + box[2]) / 2), int((box[1] + box[3]) / 2)
            out.append(Prim("arrow", (cx, cy), (int(cx + vx * self.velocity_scale), int(cy + vy * self.velocity_scale)),
                            self.colors["velocity"], 2))

    def compile_frame(self, shape: Sequence[int], tracks: Iterable[dict], detections=None, frame_info: Optional[dict] = None) -> list:
        """The frame's display list in painting order: raw detections, tracks, counters, legend (:29-44)."""
        tracks = list(tracks)
        out: list = []
        if detections is not None and len(detections):
            for det in detections:
                if len(det) >= 5:
                    x1, y1, x2, y2 = (int(v) for v in det[:4])
                    out.append(Prim("box", (x1, y1), (x2, y2), self.colors["detected"], 1))
                    out.append(Prim("text", (x1, y1 - 5), (), self.colors["detected"], 1, f"Det: {det[4]:.2f}", 0.3))
        for t in tracks:
            self._compile_track(out, t, shape)
        if frame_info:
            by_state = {k: sum(1 for t in tracks if t.get("status") == k) for k in ("detected", "predicted")}
            lines = [f"Frame: {frame_info.get('frame_number', 0)}",
                     f"Detections: {len(detections) if detections is not None and len(detections) else 0}",
                     f"Tracking (Green): {by_state['detected']}",
                     f"Predicting (Orange): {by_state['predicted']}"]
            if "state_changes" in frame_info:
                lines.append(f"State Changes: {frame_info['state_changes']}")
            for i, line in enumerate(lines):
                out.append(Prim("text", (10, 30 + 25 * i), (), self.colors["text"], 2, line, 0.6))
        # legend plate in the lower right corner (:210-234)
        h, w = shape[:2]
        lx, ly = w - 220, h - 100
        out.append(Prim("box", (lx - 10, ly - 10), (w - 10, h - 10), self.colors["background"], -1))
        out.append(Prim("box", (lx - 10, ly - 10), (w - 10, h - 10), self.colors["text"], 2))
        out.append(Prim("text", (lx, ly - 5), (), self.colors["text"], 2, "Status Legend", 0.6))
        for i, (label, key) in enumerate((("Green = Detection", "detected"), ("Orange = Prediction", "predicted"), ("Yellow = Trail", "trajectory"))):
            y = ly + 15 + 20 * i
            out.append(Prim("box", (lx, y), (lx + 15, y + 15), self.colors[key], -1))
            out.append(Prim("text", (lx + 25, y + 12), (), self.colors["text"], 1, label, 0.45))
        return out

    # ---- rasteriser ---------------------------------------------------------------------------------------------
    @staticmethod
    def rasterize(image: np.ndarray, prims: Iterable[Prim]) -> np.ndarray:
        """Execute a display list on `image` in place."""
        for p in prims:
            if p.kind == "box":
                cv2.rectangle(image, p.a, p.b, p.color, p.thickness)
            elif p.kind == "tint":
                wash = image.copy()
                cv2.rectangle(wash, p.a, p.b, p.color, -1)
                cv2.addWeighted(wash, 0.3, image, 0.7, 0, image)
            elif p.kind == "text":
                cv2.putText(image, p.text, p.a, _FONT, p.scale, p.color, p.thickness)
            elif p.kind == "seg":
                cv2.line(image, p.a, p.b, p.color, p.thickness)
            elif p.kind == "arrow":
                cv2.arrowedLine(image, p.a, p.b, p.color, p.thickness, tipLength=0.3)
            else:
                raise ValueError(f"unknown primitive {p.kind!r}")
        return image

    def draw_tracks(self, image: np.ndarray, tracks, detections=None, frame_info: Optional[dict] = None) -> np.ndarray:
        """Annotated copy of `image` (the input is left untouched); advances the blink counter once per call."""
        self.frame_counter += 1
        return self.rasterize(image.copy(), self.compile_frame(image.shape, tracks, detections, frame_info))
