"""Frame input for the offline runner: the `*.png` + `*_depth.npy` layout of the reference's vo_stereo_runner.py
(:38-50), decoded ahead of the consumer by worker threads so that PNG decoding and `.npy` reads overlap the GPU work
of the previous frames.  Order and contents are exactly those of the reference's sequential loop:
sorted glob of both patterns, zipped; BGR -> RGB; depth as stored.
"""
import glob
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np


def list_frames(img_path):
    """[(png, depth_npy)] in the reference's order (sorted globs zipped, vo_stereo_runner.py:38-41)."""
    images = sorted(glob.glob(img_path + "/*.png"))
    depths = sorted(glob.glob(img_path + "/*_depth.npy"))
    return list(zip(images, depths))


def load_frame(image_file, depth_file):
    frame = cv2.imread(image_file)
    if frame is None:
        raise FileNotFoundError(image_file)
    depth = np.load(depth_file)
    return cv2.cvtColor(frame, cv2.COLOR_BGR2RGB), depth


class FramePrefetcher:
    """Iterates (index, rgb, depth) with up to `ahead` frames being decoded in `workers` threads."""

    def __init__(self, img_path, ahead=4, workers=4):
        self.files = list_frames(img_path)
        self.ahead, self.workers = max(1, int(ahead)), max(1, int(workers))

    def __len__(self):
        return len(self.files)

    def __iter__(self):
        pending = deque()
        with ThreadPoolExecutor(max_workers=self.workers) as pool:
            it = iter(enumerate(self.files))
            for index, (png, npy) in it:
                pending.append((index, pool.submit(load_frame, png, npy)))
                if len(pending) >= self.ahead:
                    break
            while pending:
                index, fut = pending.popleft()
                nxt = next(it, None)
                if nxt is not None:
                    pending.append((nxt[0], pool.submit(load_frame, *nxt[1])))
                rgb, depth = fut.result()          # re-raises a worker's exception at the right frame
                yield index, rgb, depth
