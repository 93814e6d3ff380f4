#!/usr/bin/env python
"""Golden FIRST-PERSON frames (--visible_radius > 0) from the real OpenCV, following the reference's call sequence:
XMap::to_image (xmap.cpp:125-205): white canvas, per-item XItem::get_item_image (xitem.cpp:33-63: getRotationMatrix2D +
warpAffine with the item's yaw / scale / offset, white border), image_masking (ROI ahead of the agent + wall shadows),
copyMakeBorder (black), crop, black shadow cells, warpAffine by 90 + yaw about the view centre -> get_screen_rgb
(xworld_simulator.cpp:287-307): cv::resize to the map's pixel size, HWC -> CHW -> down_sample_image (:508-545): CHW -> HWC,
cv::resize to the frame, HWC -> CHW.

Pixel operations are the real cv2's; the ROI and the shadow flags come from the reference's own XMap::image_masking,
compiled from xmap.cpp where it lies (oracle/_ref, ref_map_masking).

Runs only in the build container (needs /root/reference's icons, cv2 and oracle/_ref).
Output: tests/golden/fpv_golden.npz.
"""
import ctypes as C
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from xworld_b200.catalog import Catalog  # noqa: E402

ITEM_PATH = "/root/reference/games/xworld/images"
cv2.setNumThreads(1)
cv2.ipp.setUseIPP(False)  # the reference builds OpenCV WITH_IPP=OFF (cmake/opencv.cmake:22)
PI_2 = 1.5707963


def ref_lib():
    R = C.CDLL(oracle.REF_LIB)
    R.ref_map_create.restype = C.c_void_p
    R.ref_map_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
    R.ref_map_destroy.argtypes = [C.c_void_p]
    R.ref_map_masking.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return R


def get_item_image(icon, yaw, scale, offset):
    """xitem.cpp:47-60"""
    icon = icon.copy()
    center = (icon.shape[1] / 2.0, icon.shape[0] / 2.0)
    rot = cv2.getRotationMatrix2D(center, 90 - yaw * 180 / math.pi, scale)
    rot[0, 2] += (offset + scale / 2 - 0.5) * icon.shape[1]
    rot[1, 2] += (offset + scale / 2 - 0.5) * icon.shape[0]
    return cv2.warpAffine(icon, rot, (icon.shape[1], icon.shape[0]), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=(255, 255, 255))


def masking(R, grid, H, W, agent_yaw, vr):
    cells = np.nonzero(grid)[0]
    types = np.array([0 if grid[c] == 1 else 2 if grid[c] == 2 else 1 for c in cells], np.int32)
    xs, ys = (cells % W).astype(np.int32), (cells // W).astype(np.int32)
    m = R.ref_map_create(H, W, len(cells), types.ctypes.data, xs.ctypes.data, ys.ctypes.data, agent_yaw, vr)
    rect = np.zeros(4, np.int32)
    shadow = np.zeros(vr * vr, np.uint8)
    R.ref_map_masking(m, vr, rect.ctypes.data, shadow.ctypes.data)
    R.ref_map_destroy(m)
    return rect, shadow


def render(R, grid, H, W, vr, agent_yaw, goal_icon, goal_pose, atlas, brick, robot, oh, ow):
    world = np.full((H * 64, W * 64, 3), 255, np.uint8)
    for i in range(H):
        for j in range(W):
            code = grid[i * W + j]
            if code == 0:
                continue
            if code == 1:
                img = get_item_image(atlas[brick], PI_2, 1.0, 0.0)
            elif code == 2:
                img = get_item_image(atlas[robot], agent_yaw, 1.0, 0.0)
            else:
                y, s, o = goal_pose[code - 3]
                img = get_item_image(atlas[goal_icon[code - 3]], y, s, o)
            world[i * 64:(i + 1) * 64, j * 64:(j + 1) * 64] = img
    rect, shadow = masking(R, grid, H, W, agent_yaw, vr)
    assert rect[2] == vr and rect[3] == vr
    world = cv2.copyMakeBorder(world, vr * 64, vr * 64, vr * 64, vr * 64, cv2.BORDER_CONSTANT, value=(0, 0, 0))
    view = world[rect[1] * 64:(rect[1] + vr) * 64, rect[0] * 64:(rect[0] + vr) * 64].copy()
    for x in range(vr):
        for y in range(vr):
            if shadow[y * vr + x]:
                view[y * 64:(y + 1) * 64, x * 64:(x + 1) * 64] = 0
    rot = cv2.getRotationMatrix2D((view.shape[1] / 2.0, view.shape[0] / 2.0), 90 + agent_yaw * 180 / math.pi, 1.0)
    view = cv2.warpAffine(view, rot, (view.shape[1], view.shape[0]))
    screen = cv2.resize(view, (W * 64, H * 64), interpolation=cv2.INTER_LINEAR)        # get_screen_rgb
    planar = np.ascontiguousarray(screen.transpose(2, 0, 1))
    img = np.ascontiguousarray(planar.transpose(1, 2, 0))                                 # down_sample_image
    out = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def turned(yaw, turns):
    """XAgent::act TURN_LEFT (< 0) / TURN_RIGHT (> 0) applied to a yaw (xitem.cpp:140-151): the drifting doubles."""
    for _ in range(abs(turns)):
        if turns > 0:
            yaw += math.pi / 2
            if yaw > math.pi + 1e-4:
                yaw -= 2 * math.pi
        else:
            yaw -= math.pi / 2
            if yaw < -math.pi / 2 - 1e-4:
                yaw += 2 * math.pi
    return yaw


CASES = (("vr3_7x7", 7, 3, 12), ("vr7_11x11", 11, 7, 30), ("vr5_8x8", 8, 5, 16), ("vr7_7x7", 7, 7, 12), ("vr1_7x7", 7, 1, 12),
         ("vr9_15x15", 15, 9, 56), ("vr11_11x11", 11, 11, 30))


def main():
    R = ref_lib()
    cat = Catalog.from_item_path(ITEM_PATH)
    keep = [cat.brick_icon, cat.agent_icon] + [int(i) for i in cat.name_icons[::26][:14]]
    atlas = cat.atlas64[keep]                      # 16 icons; 0 = brick, 1 = robot, 2.. = goals
    rng = np.random.RandomState(11)
    out = {"atlas": atlas, "paths": np.array([cat.icon_meta[i]["path"] for i in keep])}
    for tag, H, vr, n_block in CASES:
        block = 84 // vr
        oh = vr * block
        grids, icons, poses, yaws, frames = [], [], [], [], []
        for t in range(6):
            cells = rng.permutation(H * H)[:n_block + 5]
            grid = np.zeros(H * H, np.uint8)
            grid[cells[:n_block]] = 1
            grid[cells[n_block]] = 2
            if t == 0:  # the agent in a corner: the window hangs over the black border
                grid[grid == 2] = 0
                grid[0] = 2
            gi = rng.permutation(14)[:4] + 2
            pose = []
            for k in range(4):
                if grid[cells[n_block + 1 + k]] == 0:
                    grid[cells[n_block + 1 + k]] = 3 + k
                yaw = PI_2 * 4 * (rng.randint(0, 4096) / 4096.0)   # the engine's yaw grid (oracle/xw_oracle_fpv.c xo_goal_pose)
                scale = 0.5 + 0.5 * rng.rand()
                pose.append((yaw, scale, (1 - scale) * rng.rand()))
            agent_yaw = turned([-1, 0, 1, 2][t % 4] * PI_2, [0, 0, 0, 0, 5, -7][t])
            grids.append(grid)
            icons.append(gi.astype(np.int32))
            poses.append(np.array(pose, np.float64))
            yaws.append(agent_yaw)
            frames.append(render(R, grid, H, H, vr, agent_yaw, gi, pose, atlas, 0, 1, oh, oh))
        out[tag + "_grid"] = np.stack(grids)
        out[tag + "_goal_icon"] = np.stack(icons)
        out[tag + "_goal_pose"] = np.stack(poses)
        out[tag + "_agent_yaw"] = np.array(yaws, np.float64)
        out[tag + "_frames"] = np.stack(frames)
        print(tag, "frames", out[tag + "_frames"].shape)
    path = os.path.join(HERE, "fpv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
