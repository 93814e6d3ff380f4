"""Icon catalog: what XWorldEnv builds from `item_path` (games/xworld/maps/xworld_env.py:76-94,
set_goal_subtrees :244-268) -- the goal class names, their icon variants, the colour table -- plus
the decoded 64x64 BGR atlas the renderer composites from (XItem::get_item_image, xitem.cpp:33-43).

Two sources:
  * Catalog.from_item_path(dir): the reference's own images directory, decoded with cv2.imread
    (flag 1) exactly as the reference does; parity is defined post-decode.
  * Catalog.synthetic(): same names / variants / colours as the reference's directory (metadata in
    assets/xworld_icons.json), procedurally generated pixels.  Used by bench.py and the GPU tests,
    which run on a box without /root/reference.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _abi

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
NAV_SUBTREES = ("animal", "fruit", "furniture", "vegetable")  # XWorldNav.py:17


class Catalog(object):
    def __init__(self, icon_meta, atlas64, subtrees=NAV_SUBTREES):
        """icon_meta: list of dicts {path,type,name,subtree,color} sorted by path;
        atlas64: uint8 [n_icons,64,64,3] BGR."""
        self.icon_meta = list(icon_meta)
        self.atlas64 = np.ascontiguousarray(atlas64, dtype=np.uint8)
        assert self.atlas64.shape == (len(self.icon_meta), 64, 64, 3)
        self.subtrees = tuple(subtrees)
        bricks = [i for i, m in enumerate(self.icon_meta) if m["type"] == "block" and m["name"] == "brick"]
        agents = [i for i, m in enumerate(self.icon_meta) if m["type"] == "agent"]
        assert bricks and agents, "catalog needs block/brick_* and agent/* icons"
        self.brick_icon, self.agent_icon = bricks[0], agents[0]
        # goal names of the selected subtrees, canonical = sorted (xworld_env.py:256-268)
        by_name = {}
        for i, m in enumerate(self.icon_meta):
            if m["type"] == "goal" and (not self.subtrees or m["subtree"] in self.subtrees):
                by_name.setdefault(m["name"], []).append(i)
        self.names = sorted(by_name)
        first, icons = [0], []
        for n in self.names:
            icons.extend(sorted(by_name[n], key=lambda i: self.icon_meta[i]["path"]))
            first.append(len(icons))
        self.name_first = np.asarray(first, dtype=np.int32)
        self.name_icons = np.asarray(icons, dtype=np.int32)
        self.icon_colored = np.asarray([m["color"] != "na" for m in self.icon_meta], dtype=np.uint8)
        self._c = None

    @property
    def n_icons(self):
        return len(self.icon_meta)

    def as_c(self):
        """xw_catalog struct (keeps the numpy buffers alive through self)."""
        if self._c is None:
            c = _abi.XwCatalog()
            c.n_icons = self.n_icons
            c.brick_icon, c.agent_icon = self.brick_icon, self.agent_icon
            c.n_names = len(self.names)
            c.name_first = self.name_first.ctypes.data_as(C.POINTER(C.c_int32))
            c.name_icons = self.name_icons.ctypes.data_as(C.POINTER(C.c_int32))
            c.icon_colored = self.icon_colored.ctypes.data_as(C.POINTER(C.c_uint8))
            c.atlas64 = self.atlas64.ctypes.data_as(C.POINTER(C.c_uint8))
            self._c = c
        return self._c

    # ------------------------------------------------------------------ sources
    @staticmethod
    def metadata():
        with open(os.path.join(_ASSETS, "xworld_icons.json")) as f:
            return json.load(f)["icons"]

    @classmethod
    def from_item_path(cls, item_path, subtrees=NAV_SUBTREES):
        import cv2
        metas = []
        colors = {}
        prop = os.path.join(item_path, "properties.txt")
        with open(prop) as f:
            for l in f.read().splitlines():
                if l.startswith("//") or l == "":
                    continue
                colors[l.split()[0]] = l.split()[1]
        for dp, _, fs in os.walk(item_path):
            for fn in fs:
                if fn.endswith(".jpg") or fn.endswith(".png"):
                    rel = os.path.relpath(os.path.join(dp, fn), item_path)
                    parts = rel.split(os.sep)
                    typ = [t for t in parts if t in ("goal", "block", "agent")]
                    if not typ:
                        continue
                    metas.append({
                        "path": rel, "type": typ[0],
                        "name": "_".join(os.path.basename(rel).split("_")[:-1]),
                        "subtree": parts[-2] if typ[0] == "goal" else "",
                        "color": colors.get(rel, "na")})
        metas.sort(key=lambda m: m["path"])
        atlas = np.zeros((len(metas), 64, 64, 3), np.uint8)
        for i, m in enumerate(metas):
            img = cv2.imread(os.path.join(item_path, m["path"]), 1)
            if img is None:
                raise RuntimeError("could not open or find the image: " + m["path"])
            if img.shape[:2] != (64, 64):
                img = cv2.resize(img, (64, 64), interpolation=cv2.INTER_LINEAR)  # xitem.cpp:40
            atlas[i] = img
        return cls(metas, atlas, subtrees)

    @classmethod
    def synthetic(cls, seed=0, subtrees=NAV_SUBTREES, max_icons=None):
        """Reference catalog structure, procedural pixels (deterministic in `seed`)."""
        metas = cls.metadata()
        if max_icons is not None:  # keep brick + agent + the first goals
            keep = [m for m in metas if m["type"] != "goal"]
            goals = [m for m in metas if m["type"] == "goal" and m["subtree"] in subtrees]
            metas = sorted(keep + goals[:max_icons], key=lambda m: m["path"])
        n = len(metas)
        rng = np.random.RandomState(seed)
        yy, xx = np.mgrid[0:64, 0:64].astype(np.float32)
        atlas = np.empty((n, 64, 64, 3), np.uint8)
        for i in range(n):
            col = rng.randint(0, 256, size=(2, 3)).astype(np.float32)
            fx, fy, ph = rng.uniform(0.05, 0.6), rng.uniform(0.05, 0.6), rng.uniform(0, 6.28)
            w = 0.5 + 0.5 * np.sin(xx * fx + yy * fy + ph)
            r = np.hypot(xx - 31.5, yy - 31.5)
            img = col[0][None, None, :] * w[..., None] + col[1][None, None, :] * (1 - w[..., None])
            img[r > rng.uniform(20, 34)] = 255.0  # white surround like the real icons
            img += rng.uniform(-6, 6, size=img.shape)
            atlas[i] = np.clip(img, 0, 255).astype(np.uint8)
        return cls(metas, atlas, subtrees)
