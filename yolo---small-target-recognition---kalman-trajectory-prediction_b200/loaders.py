"""Frame ingest for file sources: mirror of ``ultralytics/data/loaders.py`` ``LoadImagesAndVideos`` (:346-490) for image files,
video files, directories, globs and ``.txt`` / ``.csv`` lists (``IMG_FORMATS`` / ``VID_FORMATS``: data/utils.py:39-40).

Host-side plumbing in front of the hot path (OpenCV decode, as in the reference): it hands batches of HWC BGR uint8 frames to
``YOLO.predict`` / ``YOLO.track``; letterbox, colour conversion and normalisation happen inside the stem kernel on the GPU.
"""
from __future__ import annotations

import glob
import math
import os
from pathlib import Path

import numpy as np

IMG_FORMATS = {"bmp", "dng", "jpeg", "jpg", "mpo", "png", "tif", "tiff", "webp", "pfm"}        # heic needs pi-heif: not provided
VID_FORMATS = {"asf", "avi", "gif", "m4v", "mkv", "mov", "mp4", "mpeg", "mpg", "ts", "wmv", "webm"}
FORMATS_HELP_MSG = f"Supported formats are:\nimages: {IMG_FORMATS}\nvideos: {VID_FORMATS}"


class LoadImagesAndVideos:
    """Iterates ``(paths, imgs, info)`` batches of up to ``batch`` frames; ``vid_stride`` skips video frames (grab without retrieve).
    ``mode`` is 'image' or 'video' (what the reference's predictor uses to decide between one tracker per video)."""

    def __init__(self, path, batch=1, vid_stride=1, channels=3):
        parent = None
        if isinstance(path, (str, Path)) and Path(path).suffix in {".txt", ".csv"}:
            parent, content = Path(path).parent, Path(path).read_text()
            path = [p.strip() for p in (content.splitlines() if Path(path).suffix == ".txt" else content.split(","))]
        files = []
        for p in sorted(path) if isinstance(path, (list, tuple)) else [path]:
            a = str(Path(p).absolute())
            if "*" in a:
                files.extend(sorted(glob.glob(a, recursive=True)))
            elif os.path.isdir(a):
                files.extend(sorted(glob.glob(os.path.join(a, "*.*"))))
            elif os.path.isfile(a):
                files.append(a)
            elif parent and (parent / p).is_file():
                files.append(str((parent / p).absolute()))
            else:
                raise FileNotFoundError(f"{p} does not exist")
        images = [f for f in files if f.rpartition(".")[-1].lower() in IMG_FORMATS]
        videos = [f for f in files if f.rpartition(".")[-1].lower() in VID_FORMATS]
        self.files, self.ni, self.nf = images + videos, len(images), len(images) + len(videos)
        self.video_flag = [False] * len(images) + [True] * len(videos)
        self.mode = "video" if not images else "image"
        self.vid_stride, self.bs, self.channels = vid_stride, batch, channels
        self.cap = None
        if self.nf == 0:
            raise FileNotFoundError(f"No images or videos found in {path}. {FORMATS_HELP_MSG}")
        if videos:
            self._new_video(videos[0])

    def __len__(self):
        return math.ceil(self.nf / self.bs)

    def __iter__(self):
        self.count = 0
        return self

    def _new_video(self, path):
        import cv2

        self.frame = 0
        self.cap = cv2.VideoCapture(path)
        self.fps = int(self.cap.get(cv2.CAP_PROP_FPS))
        if not self.cap.isOpened():
            raise FileNotFoundError(f"Failed to open video {path}")
        self.frames = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT) / self.vid_stride)

    def __next__(self):
        import cv2

        paths, imgs, info = [], [], []
        while len(imgs) < self.bs:
            if self.count >= self.nf:
                if imgs:
                    return paths, imgs, info
                raise StopIteration
            path = self.files[self.count]
            if self.video_flag[self.count]:
                self.mode = "video"
                if not self.cap or not self.cap.isOpened():
                    self._new_video(path)
                success = False
                for _ in range(self.vid_stride):
                    success = self.cap.grab()
                    if not success:
                        break
                if success:
                    success, im0 = self.cap.retrieve()
                    if success:
                        if self.channels == 1:
                            im0 = cv2.cvtColor(im0, cv2.COLOR_BGR2GRAY)[..., None]
                        self.frame += 1
                        paths.append(path); imgs.append(im0)
                        info.append(f"video {self.count + 1}/{self.nf} (frame {self.frame}/{self.frames}) {path}: ")
                        if self.frame == self.frames:
                            self.count += 1
                            self.cap.release()
                else:
                    self.count += 1
                    if self.cap:
                        self.cap.release()
                    if self.count < self.nf:
                        self._new_video(self.files[self.count])
            else:
                self.mode = "image"
                im0 = cv2.imdecode(np.fromfile(path, np.uint8), cv2.IMREAD_GRAYSCALE if self.channels == 1 else cv2.IMREAD_COLOR)   # utils/patches.py imread
                if im0 is not None:
                    paths.append(path); imgs.append(im0 if im0.ndim == 3 else im0[..., None])
                    info.append(f"image {self.count + 1}/{self.nf} {path}: ")
                self.count += 1
                if self.count >= self.ni:
                    break
        return paths, imgs, info
