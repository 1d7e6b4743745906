"""Headless presenter (reference: rendering/_presentation.py).

The reference opens an SDL2 window (prebuilt Windows DLLs, _presentation.py:5-8, 24-37) and copies the
BGRA8 render target into a streaming texture every present().  On a B200 server there is no display, so this
Presenter owns the render target (same create_image2d(w, h, RGBA), :39) and counts frames instead:

  RENDERTOY_B200_FRAMES   frames to present before poll_events() reports Event.CLOSED (default 1), so the
                          tutorials' `while True:` loops terminate
  RENDERTOY_B200_DUMP     directory to write frame_NNNN.png into at each present() (BGRA -> RGB as in
                          Class2022/.../scene.py:74-79)
"""
import os
import time
from enum import Enum

import numpy as np

from ._core import create_image2d, RGBA, mapped


class Event(Enum):
    NONE = 0
    CLOSED = 1


class Presenter:
    def __init__(self, width, height, offline: bool):
        self.width = width
        self.height = height
        self.offline = offline
        self.window = None
        self.render_target = create_image2d(self.width, self.height, RGBA)
        self.first_time = None
        self.frames = 0
        self.frame_budget = int(os.environ.get("RENDERTOY_B200_FRAMES", "1"))
        self.dump_dir = os.environ.get("RENDERTOY_B200_DUMP")
        self.fps = 0.0

    def get_render_target(self):
        return self.render_target

    def is_alive(self):
        return self.frames < self.frame_budget

    def poll_events(self):
        if self.offline or self.frames >= self.frame_budget:
            return Event.CLOSED, None
        return Event.NONE, None

    def _copy_render_target(self, buffer):
        with mapped(self.render_target) as map:
            buffer[:] = map.ravel()

    def present(self):
        if self.offline:
            return
        if self.dump_dir:
            from PIL import Image as PILImage
            os.makedirs(self.dump_dir, exist_ok=True)
            bgra = self.render_target.get()
            PILImage.fromarray(np.ascontiguousarray(bgra[:, :, [2, 1, 0]])).save(
                os.path.join(self.dump_dir, f"frame_{self.frames:04d}.png"))
        if self.first_time is None:
            self.first_time = time.perf_counter()
        self.frames += 1
        self.fps = self.frames / max(0.00000001, time.perf_counter() - self.first_time)


def create_presenter(width: int, height: int) -> Presenter:
    return Presenter(width, height, False)
