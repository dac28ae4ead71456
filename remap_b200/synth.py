"""Seeded synthetic frame sequences for the registration path.

Frames are what ``nil::read_raw`` hands the reference (src/nil.hpp:14-31): ``H*W`` bytes per frame,
row-major, one C64 colour code 0..15 per byte.  The generator draws an 8x8-tile world with detail
at two scales plus speckle (flat or 2-colour checker tiles give no keypoints at all, SURVEY.md
Appendix C), then scrolls a ``W x H`` camera window over it with a piecewise-constant velocity.
Optional extras: moving sprites (foreground), hard scene cuts, a second parallax layer.

Pure numpy; used by tests, bench.py and the oracle fixtures.  Data generation only -- no part of
the registration path lives here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

TILE = 8


@dataclass
class Sequence:
    frames: np.ndarray            # (N, H, W) uint8, values 0..15
    path: np.ndarray              # (N, 2) int32 camera (x, y) of the top-left corner
    level: np.ndarray             # (N,) int32 world/level index (changes at scene cuts)
    meta: dict = field(default_factory=dict)

    @property
    def true_offsets(self) -> np.ndarray:
        """(N-1, 2) ground-truth ``prev - curr`` keypoint offsets == -(camera delta).

        kpm reports ``prev.xy - curr.xy`` (src/kpm.hpp:96-98): content that moved left by d on
        screen (camera moved right by d) gives dx = +d.
        """
        return (self.path[1:] - self.path[:-1]).astype(np.int32)


def make_world(rng: np.random.Generator, world_w: int, world_h: int, n_tiles: int = 64,
               speckle: float = 0.05, detail: int = 1) -> np.ndarray:
    """World map (world_h, world_w) uint8 in 0..15 built from ``n_tiles`` random 8x8 tiles."""
    tiles = np.empty((n_tiles, TILE, TILE), np.uint8)
    for t in range(n_tiles):
        base = rng.integers(0, 16, size=(2, 2), dtype=np.uint8)          # 4x4 quadrants
        tile = np.kron(base, np.ones((4, 4), np.uint8))
        nblk = int(rng.integers(0, 2 * detail + 1))                      # a few 2x2 .. 3x3 blobs
        for _ in range(nblk):
            bw, bh = int(rng.integers(2, 4)), int(rng.integers(2, 4))
            bx, by = int(rng.integers(0, TILE - bw + 1)), int(rng.integers(0, TILE - bh + 1))
            tile[by:by + bh, bx:bx + bw] = rng.integers(0, 16, dtype=np.uint8)
        nline = int(rng.integers(0, detail))                             # thin lines (detail >= 2)
        for _ in range(nline):
            if rng.integers(0, 2):
                tile[int(rng.integers(0, TILE)), :] = rng.integers(0, 16, dtype=np.uint8)
            else:
                tile[:, int(rng.integers(0, TILE))] = rng.integers(0, 16, dtype=np.uint8)
        tiles[t] = tile
    ty, tx = world_h // TILE, world_w // TILE
    layout = rng.integers(0, n_tiles, size=(ty, tx))
    world = tiles[layout].transpose(0, 2, 1, 3).reshape(ty * TILE, tx * TILE)
    if speckle > 0:
        m = rng.random(world.shape) < speckle
        world = world.copy()
        world[m] = rng.integers(0, 16, size=int(m.sum()), dtype=np.uint8)
    return np.ascontiguousarray(world)


def camera_path(rng: np.random.Generator, n: int, w: int, h: int, world_w: int, world_h: int,
                vmax=(4, 3), seg=(10, 60)) -> np.ndarray:
    """Piecewise-constant velocity camera path, reflected at the world border."""
    path = np.empty((n, 2), np.int32)
    x = int(rng.integers(0, world_w - w + 1))
    y = int(rng.integers(0, world_h - h + 1))
    i = 0
    while i < n:
        vx = int(rng.integers(-vmax[0], vmax[0] + 1))
        vy = int(rng.integers(-vmax[1], vmax[1] + 1))
        length = int(rng.integers(seg[0], seg[1] + 1))
        for _ in range(length):
            if i >= n:
                break
            path[i] = (x, y)
            i += 1
            nx, ny = x + vx, y + vy
            if nx < 0 or nx > world_w - w:
                vx = -vx
                nx = x + vx
            if ny < 0 or ny > world_h - h:
                vy = -vy
                ny = y + vy
            x, y = nx, ny
    return path


def render(world: np.ndarray, path: np.ndarray, w: int, h: int) -> np.ndarray:
    n = path.shape[0]
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        x, y = int(path[i, 0]), int(path[i, 1])
        out[i] = world[y:y + h, x:x + w]
    return out


class TilemapPlan:
    """Everything about a scrolling_tilemap sequence except the pixels: worlds, camera path, levels, sprite set.
    render(a, b) draws frames [a, b); building the plan once lets a long sequence be rendered batch by batch
    (tools/run_config.py) with the very same result as one scrolling_tilemap call."""

    def __init__(self, n: int, w: int = 320, h: int = 224, seed: int = 1, *, world_w: int = 4096, world_h: int = 2048,
                 n_tiles: int = 64, speckle: float = 0.05, vmax=(4, 3), sprites: int = 0, cut_every: int = 0,
                 levels: int = 1, parallax: int = 0, detail: int = 1, sprite_motion: str = "stateful"):
        rng = np.random.default_rng(seed)
        self.n, self.w, self.h = n, w, h
        world_w = max(world_w, w + 64)
        world_h = max(world_h, h + 64)
        self.worlds = [make_world(rng, world_w, world_h, n_tiles, speckle, detail) for _ in range(max(1, levels))]
        path = camera_path(rng, n, w, h, world_w, world_h, vmax)
        level = np.zeros(n, np.int32)
        if cut_every > 0:
            i = int(rng.integers(cut_every // 2 + 1, cut_every * 3 // 2 + 2))
            cur = 0
            while i < n:
                cur = (cur + 1) % len(self.worlds) if len(self.worlds) > 1 else cur
                jump = np.array([int(rng.integers(0, world_w - w + 1)), int(rng.integers(0, world_h - h + 1))])
                delta = jump - path[i]
                path[i:] += delta                                   # teleport, keep the velocity plan
                np.clip(path[i:, 0], 0, world_w - w, out=path[i:, 0])
                np.clip(path[i:, 1], 0, world_h - h, out=path[i:, 1])
                level[i:] = cur
                i += int(rng.integers(cut_every // 2 + 1, cut_every * 3 // 2 + 2))
        self.path, self.level = path, level
        self.parallax = parallax
        self.far = make_world(rng, world_w, world_h, n_tiles, speckle, detail) if parallax > 0 else None
        self.sprites, self.sprite_motion = sprites, sprite_motion
        if sprites > 0:
            self.sw = rng.integers(12, 25, size=sprites)
            self.sh = rng.integers(12, 25, size=sprites)
            self.tex = [rng.integers(0, 16, size=(int(self.sh[k]), int(self.sw[k])), dtype=np.uint8) for k in range(sprites)]
            self.pos0 = np.stack([rng.integers(0, w - 24, size=sprites), rng.integers(0, h - 24, size=sprites)], 1).astype(np.float64)
            self.vel0 = rng.uniform(-3, 3, size=(sprites, 2))
        self.meta = dict(n=n, w=w, h=h, seed=seed, n_tiles=n_tiles, speckle=speckle, vmax=tuple(vmax), sprites=sprites,
                         cut_every=cut_every, levels=levels, parallax=parallax, detail=detail, sprite_motion=sprite_motion)

    def render(self, a: int = 0, b: int | None = None, out: np.ndarray | None = None) -> Sequence:
        b = self.n if b is None else b
        n, w, h = b - a, self.w, self.h
        assert self.sprites == 0 or self.sprite_motion == "closed" or a == 0, \
            "stateful sprites need the sequence rendered from frame 0; use sprite_motion='closed'"
        path, level = self.path[a:b].copy(), self.level[a:b].copy()
        frames = out if out is not None else np.empty((n, h, w), np.uint8)
        assert frames.shape == (n, h, w) and frames.dtype == np.uint8
        for i in range(n):
            x, y = int(path[i, 0]), int(path[i, 1])
            frames[i] = self.worlds[level[i]][y:y + h, x:x + w]
        if self.parallax > 0:
            band = (np.arange(h) // self.parallax) % 2 == 1
            for i in range(n):
                x, y = int(path[i, 0]) // 2, int(path[i, 1]) // 2
                frames[i][band] = self.far[y:y + h, x:x + w][band]
        if self.sprites > 0:
            sw, sh, tex, sprites = self.sw, self.sh, self.tex, self.sprites
            if self.sprite_motion == "closed":
                span = np.stack([w - sw, h - sh], 1).astype(np.float64)          # a sprite stays inside the frame
                for i in range(n):
                    p = self.pos0 + self.vel0 * float(a + i)
                    m = np.mod(p, 2 * span)
                    p = np.where(m > span, 2 * span - m, m)
                    for k in range(sprites):
                        px, py = int(p[k, 0]), int(p[k, 1])
                        frames[i, py:py + int(sh[k]), px:px + int(sw[k])] = tex[k][:h - py, :w - px]
            else:
                pos, vel = self.pos0.copy(), self.vel0.copy()
                for i in range(n):
                    for k in range(sprites):
                        px, py = int(pos[k, 0]), int(pos[k, 1])
                        frames[i, py:py + int(sh[k]), px:px + int(sw[k])] = tex[k][:h - py, :w - px]
                    pos += vel
                    for k in range(sprites):
                        if pos[k, 0] < 0 or pos[k, 0] > w - sw[k]:
                            vel[k, 0] = -vel[k, 0]
                            pos[k, 0] = min(max(pos[k, 0], 0), w - sw[k])
                        if pos[k, 1] < 0 or pos[k, 1] > h - sh[k]:
                            vel[k, 1] = -vel[k, 1]
                            pos[k, 1] = min(max(pos[k, 1], 0), h - sh[k])
        return Sequence(frames=frames, path=path, level=level, meta=dict(self.meta, n=n))


def scrolling_tilemap(n: int, w: int = 320, h: int = 224, seed: int = 1, *, world_w: int = 4096,
                      world_h: int = 2048, n_tiles: int = 64, speckle: float = 0.05,
                      vmax=(4, 3), sprites: int = 0, cut_every: int = 0, levels: int = 1,
                      parallax: int = 0, detail: int = 1, frame_range=None, sprite_motion: str = "stateful") -> Sequence:
    """The workload family of BASELINE.json ``configs``.

    sprites    number of moving textured rectangles drawn over the background (config 3)
    cut_every  mean number of frames between hard scene cuts (config 5); 0 = none
    levels     number of distinct worlds (tile sets) the cuts rotate through
    parallax   band height in pixels of a second layer scrolling at half speed; 0 = none
    frame_range (a, b): render only frames [a, b) of the n-frame sequence (a rank's shard); path and
               level still describe the rendered frames only
    sprite_motion "stateful": sprites bounce step by step (needs the whole sequence rendered in order);
               "closed": a sprite's position is a closed-form triangle wave of the frame index, so any frame_range of a
               long sequence can be rendered on its own (multi-GPU shards of config 3)
    """
    plan = TilemapPlan(n, w, h, seed, world_w=world_w, world_h=world_h, n_tiles=n_tiles, speckle=speckle, vmax=vmax,
                       sprites=sprites, cut_every=cut_every, levels=levels, parallax=parallax, detail=detail,
                       sprite_motion=sprite_motion)
    a, b = frame_range if frame_range is not None else (0, n)
    return plan.render(a, b)


def random_frames(n: int, w: int, h: int, seed: int = 0, palette: int = 16) -> np.ndarray:
    """Uniform random frames (worst case keypoint density; used by the stress tests)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, palette, size=(n, h, w), dtype=np.uint8)
