"""Frame ingest for file sources, behaviour of ``ultralytics/data/loaders.py`` ``LoadImagesAndVideos`` (:346-490): image files,
video files, directories, globs and ``.txt`` / ``.csv`` lists (suffix sets: data/utils.py:39-40), images before videos, batches of
``batch`` frames where the end of the image list closes a batch, ``vid_stride`` frames grabbed per frame returned.

Host-side plumbing in front of the hot path (OpenCV decode, as in the reference): it hands HWC BGR uint8 frames to
``YOLO.predict`` / ``YOLO.track``; letterbox, colour conversion and normalisation happen inside the stem kernel on the GPU.
Written as a frame generator plus a batcher (the reference is one ``__next__`` state machine); the batch sequence is checked
against the reference loader in tests/test_host.py.
"""
from __future__ import annotations

import glob
import math
import os
from pathlib import Path

import numpy as np

IMG_FORMATS = frozenset("bmp dng jpeg jpg mpo png tif tiff webp pfm".split())            # heic needs pi-heif: not provided
VID_FORMATS = frozenset("asf avi gif m4v mkv mov mp4 mpeg mpg ts wmv webm".split())
FORMATS_HELP_MSG = f"Supported formats are:\nimages: {sorted(IMG_FORMATS)}\nvideos: {sorted(VID_FORMATS)}"


def _suffix(name):
    return name.rpartition(".")[2].lower()


def expand_sources(path):
    """Source spec -> absolute file names in the reference's order (sorted top-level entries, globs and directories expanded)."""
    base = None
    if isinstance(path, (str, Path)) and Path(path).suffix in (".txt", ".csv"):
        listing = Path(path)
        base = listing.parent
        text = listing.read_text()
        path = [item.strip() for item in (text.splitlines() if listing.suffix == ".txt" else text.split(","))]
    names = []
    for entry in (sorted(path) if isinstance(path, (list, tuple)) else [path]):
        full = str(Path(entry).absolute())
        if "*" in full:
            names += sorted(glob.glob(full, recursive=True))
        elif os.path.isdir(full):
            names += sorted(glob.glob(os.path.join(full, "*.*")))
        elif os.path.isfile(full):
            names.append(full)
        elif base is not None and (base / entry).is_file():
            names.append(str((base / entry).absolute()))
        else:
            raise FileNotFoundError(f"{entry} does not exist")
    return names


class LoadImagesAndVideos:
    """Iterate ``(paths, imgs, info)`` batches.  ``mode`` is 'image' or 'video' (what the reference's predictor looks at to decide
    between one tracker per video and one per batch slot); ``frame`` / ``frames`` describe the video being read."""

    def __init__(self, path, batch=1, vid_stride=1, channels=3):
        names = expand_sources(path)
        stills = [n for n in names if _suffix(n) in IMG_FORMATS]
        clips = [n for n in names if _suffix(n) in VID_FORMATS]
        if not stills and not clips:
            raise FileNotFoundError(f"No images or videos found in {path}. {FORMATS_HELP_MSG}")
        self.files = stills + clips
        self.ni, self.nf = len(stills), len(stills) + len(clips)
        self.video_flag = [k >= self.ni for k in range(self.nf)]
        self.mode = "image" if stills else "video"
        self.bs, self.vid_stride, self.channels = int(batch), int(vid_stride), int(channels)
        self.cap, self.frame, self.frames, self.fps = None, 0, 0, 0
        self.count = 0
        for clip in clips:                      # the reference opens its first video at construction: a broken file fails here
            self._open(clip).release()
            break

    def __len__(self):
        return math.ceil(self.nf / self.bs)

    def _open(self, name):
        import cv2

        cap = cv2.VideoCapture(name)
        if not cap.isOpened():
            raise FileNotFoundError(f"Failed to open video {name}")
        self.fps = int(cap.get(cv2.CAP_PROP_FPS))
        self.frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT) / self.vid_stride)
        self.frame = 0
        return cap

    def _gray(self, im):
        import cv2

        return cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)[..., None] if self.channels == 1 and im.ndim == 3 and im.shape[2] == 3 else im

    def _items(self):
        """One (file index, is_last_still, path, image, info) per decoded frame, in file order."""
        import cv2

        for k, name in enumerate(self.files):
            self.count = k
            if not self.video_flag[k]:
                self.mode = "image"
                im = cv2.imdecode(np.fromfile(name, np.uint8), cv2.IMREAD_GRAYSCALE if self.channels == 1 else cv2.IMREAD_COLOR)
                if im is None:
                    continue                                  # unreadable image: skipped (the reference logs a warning)
                yield k, k == self.ni - 1, name, (im if im.ndim == 3 else im[..., None]), f"image {k + 1}/{self.nf} {name}: "
                continue
            self.mode = "video"
            self.cap = self._open(name)
            try:
                while True:
                    alive = True
                    for _ in range(self.vid_stride):          # grab vid_stride frames, decode only the last
                        alive = self.cap.grab()
                        if not alive:
                            break
                    if not alive:
                        break
                    ok, im = self.cap.retrieve()
                    if not ok:
                        continue
                    self.frame += 1
                    yield k, False, name, self._gray(im), f"video {k + 1}/{self.nf} (frame {self.frame}/{self.frames}) {name}: "
                    if self.frame == self.frames:
                        break
            finally:
                self.cap.release()
        self.count = self.nf

    def __iter__(self):
        paths, imgs, info = [], [], []
        for _, closes_batch, name, im, text in self._items():
            paths.append(name); imgs.append(im); info.append(text)
            if len(imgs) == self.bs or closes_batch:          # the end of the image list ends a batch (loaders.py:478-479)
                yield paths, imgs, info
                paths, imgs, info = [], [], []
        if imgs:
            yield paths, imgs, info


class LoadStreams:
    """Several live sources read concurrently, behaviour of ``ultralytics/data/loaders.py`` ``LoadStreams`` (:54-230): ``sources`` is
    a ``*.streams`` text file (one source per whitespace-separated token) or a single source (a video file / device index / URL that
    OpenCV can open).  One reader thread per source; every iteration hands out ONE frame per source -- the batch a multi-stream
    detect+track step consumes (``mode == 'stream'``: one tracker per source).  ``buffer=False`` keeps only the newest frame of a
    source (live cameras: drop what the consumer was too slow for), ``buffer=True`` queues up to 30 frames per source and delivers
    every frame in order.  Iteration ends when a source has run dry and its reader has finished."""

    mode = "stream"
    QUEUE = 30

    def __init__(self, sources="file.streams", vid_stride=1, buffer=False, channels=3):
        import threading

        import cv2

        self.buffer, self.vid_stride, self.channels = bool(buffer), int(vid_stride), int(channels)
        specs = Path(sources).read_text().split() if os.path.isfile(str(sources)) and str(sources).endswith(".streams") else [str(sources)]
        self.bs = len(specs)
        self.sources = ["".join(ch if ch.isalnum() or ch in "-." else "_" for ch in s) for s in specs]     # names safe for file paths
        self.running = True
        self.caps, self.queues, self.shape, self.fps, self.frames, self.threads = [], [], [], [], [], []
        self._lock = threading.Lock()
        for k, spec in enumerate(specs):
            cap = cv2.VideoCapture(int(spec) if spec.isnumeric() else spec)
            if not cap.isOpened():
                raise ConnectionError(f"{k + 1}/{self.bs}: Failed to open {spec}")
            rate = cap.get(cv2.CAP_PROP_FPS)
            total = max(int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), 0) or float("inf")                         # live sources report 0
            ok, first = cap.read()
            if not ok or first is None:
                raise ConnectionError(f"{k + 1}/{self.bs}: Failed to read images from {spec}")
            first = self._convert(first)
            self.caps.append(cap); self.queues.append([first]); self.shape.append(first.shape)
            self.fps.append(max((rate if math.isfinite(rate) else 0) % 100, 0) or 30); self.frames.append(total)
            th = threading.Thread(target=self._reader, args=(k, cap, spec), daemon=True)
            self.threads.append(th)
        for th in self.threads:
            th.start()

    def _convert(self, im):
        import cv2

        return cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)[..., None] if self.channels == 1 else im

    def _reader(self, k, cap, spec):
        import time

        seen = 0
        while self.running and cap.isOpened() and seen < self.frames[k] - 1:
            if len(self.queues[k]) >= self.QUEUE:
                time.sleep(0.01)                      # the consumer is behind: wait instead of growing the queue
                continue
            seen += 1
            cap.grab()
            if seen % self.vid_stride:
                continue
            ok, im = cap.retrieve()
            if ok:
                im = self._convert(im)
            else:                                     # signal lost: a black frame now, try to re-open the source
                im = np.zeros(self.shape[k], dtype=np.uint8)
                cap.open(int(spec) if spec.isnumeric() else spec)
            with self._lock:
                if self.buffer:
                    self.queues[k].append(im)
                else:
                    self.queues[k] = [im]

    def close(self):
        self.running = False
        for th in self.threads:
            if th.is_alive():
                th.join(timeout=5)
        for cap in self.caps:
            try:
                cap.release()
            except Exception:
                pass

    def __len__(self):
        return self.bs

    def __iter__(self):
        import time

        while True:
            batch = []
            for k in range(self.bs):
                while not self.queues[k]:
                    if not self.threads[k].is_alive():
                        if not self.queues[k]:        # (re-checked: the reader may have queued its last frame just before it ended)
                            self.close()
                            return
                        break
                    time.sleep(1 / min(self.fps))
                with self._lock:
                    q = self.queues[k]
                    if self.buffer:
                        batch.append(q.pop(0))
                    else:
                        batch.append(q[-1] if q else np.zeros(self.shape[k], dtype=np.uint8))
                        self.queues[k] = []
            yield self.sources, batch, [""] * self.bs
