#!/usr/bin/env python
"""Golden frames from the real OpenCV, following the reference's render call sequence exactly:
XMap::to_image (xmap.cpp:125-146) -> XItem::get_item_image incl. the per-frame warpAffine
(xitem.cpp:33-63) -> get_screen_rgb identity resize + HWC->CHW (xworld_simulator.cpp:287-307) ->
down_sample_image CHW->HWC, cv::resize INTER_LINEAR, HWC->CHW (:508-545).

Runs only in the build container (needs /root/reference's icons and cv2).
Output: tests/golden/render_golden.npz = 16 real decoded icons + a few maps per config + their frames.
"""
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from xworld_b200.catalog import Catalog  # noqa: E402

ITEM_PATH = "/root/reference/games/xworld/images"
cv2.setNumThreads(1)
cv2.ipp.setUseIPP(False)  # the reference builds OpenCV WITH_IPP=OFF (cmake/opencv.cmake:22)


def get_item_image(icon, yaw=1.5707963, scale=1.0, offset=0.0):
    """xitem.cpp:47-60"""
    icon = icon.copy()
    center = (icon.shape[1] / 2.0, icon.shape[0] / 2.0)
    rot = cv2.getRotationMatrix2D(center, 90 - yaw * 180 / math.pi, scale)
    rot[0, 2] += (offset + scale / 2 - 0.5) * icon.shape[1]
    rot[1, 2] += (offset + scale / 2 - 0.5) * icon.shape[0]
    return cv2.warpAffine(icon, rot, (icon.shape[1], icon.shape[0]), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=(255, 255, 255))


def render(grid, cell_icon, atlas, H, W, oh, ow):
    world = np.full((H * 64, W * 64, 3), 255, np.uint8)
    for i in range(H):
        for j in range(W):
            ic = cell_icon[i * W + j]
            if ic >= 0:
                world[i * 64:(i + 1) * 64, j * 64:(j + 1) * 64] = get_item_image(atlas[ic])
    screen = cv2.resize(world, (W * 64, H * 64), interpolation=cv2.INTER_LINEAR)  # identity
    planar = np.ascontiguousarray(screen.transpose(2, 0, 1))                         # get_screen_rgb
    img = np.ascontiguousarray(planar.transpose(1, 2, 0))                            # down_sample_image
    out = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def main():
    cat = Catalog.from_item_path(ITEM_PATH)
    keep = [cat.brick_icon, cat.agent_icon] + [int(i) for i in cat.name_icons[::26][:14]]
    atlas = cat.atlas64[keep]                      # 16 icons; 0 = brick, 1 = robot, 2.. = goals
    # identity claim of SURVEY §8a a11: warpAffine(yaw=1.5707963) returns the icon unchanged
    for i in range(cat.n_icons):
        assert (get_item_image(cat.atlas64[i]) == cat.atlas64[i]).all(), i
    rng = np.random.RandomState(7)
    out = {"atlas": atlas, "paths": np.array([cat.icon_meta[i]["path"] for i in keep])}
    for tag, H, oh in (("c2", 7, 84), ("c3", 11, 84), ("c4", 15, 128), ("ref8", 8, 96)):
        grids, icons, frames = [], [], []
        for _ in range(3):
            n_block = {7: 12, 11: 30, 15: 56, 8: 16}[H]
            cells = rng.permutation(H * H)[:n_block + 5]
            grid = np.zeros(H * H, np.uint8)
            grid[cells[:n_block]] = 1
            grid[cells[n_block]] = 2
            gi = rng.permutation(14)[:4] + 2
            for k in range(4):
                grid[cells[n_block + 1 + k]] = 3 + k
            cell_icon = np.full(H * H, -1, np.int64)
            cell_icon[grid == 1] = 0
            cell_icon[grid == 2] = 1
            for k in range(4):
                cell_icon[grid == 3 + k] = gi[k]
            grids.append(grid)
            icons.append(gi.astype(np.int32))
            frames.append(render(grid, cell_icon, atlas, H, H, oh, oh))
        out[tag + "_grid"] = np.stack(grids)
        out[tag + "_goal_icon"] = np.stack(icons)
        out[tag + "_frames"] = np.stack(frames)
    np.savez_compressed(os.path.join(HERE, "render_golden.npz"), **out)
    print("wrote render_golden.npz", os.path.getsize(os.path.join(HERE, "render_golden.npz")))


if __name__ == "__main__":
    main()
