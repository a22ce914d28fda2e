"""Headless rasteriser for ``env.render()`` with ``render_mode="rgb_array"``.

The reference draws with pygame (``ttrl_env/envs/common/graphics.py:23-230`` EnvViewer, ``road/graphics.py`` WorldSurface /
LaneGraphics / RoadGraphics, ``vehicle/graphics.py`` VehicleGraphics); pygame is not a dependency here and the GPU box has no
display.  This module redraws the same picture with numpy: the same world -> pixel transform (``scaling``, ``centering_position``,
window centred on the observer vehicle), the same lane markings (stripes of 3 m every 4.33 m, continuous lines), the same
vehicle rectangles, headlights and colours.  It is geometry-faithful, not pixel-identical (pygame's line and rotation
resampling differ in anti-aliasing); it is host-side code outside the per-step hot path.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .road import LineType, RoadNetwork

GREY, WHITE = (100, 100, 100), (255, 255, 255)                      # WorldSurface colours (road/graphics.py:24-28)
RED, GREEN, BLUE, YELLOW, BLACK = (255, 100, 100), (50, 200, 0), (100, 200, 255), (200, 200, 0), (60, 60, 60)  # vehicle/graphics.py:19-26
STRIPE_SPACING, STRIPE_LENGTH, STRIPE_WIDTH = 4.33, 3.0, 0.3         # road/graphics.py:113-124
VEH_LENGTH, VEH_WIDTH = 5.0, 2.0


class Surface:
    """``WorldSurface`` (road/graphics.py:20-104) over a numpy image."""

    def __init__(self, width: int, height: int, scaling: float, centering_position) -> None:
        self.img = np.empty((height, width, 3), np.uint8)
        self.img[:] = GREY
        self.scaling, self.centering = float(scaling), tuple(centering_position)
        self.origin = np.zeros(2)

    def move_display_window_to(self, position) -> None:
        h, w = self.img.shape[:2]
        self.origin = np.asarray(position, np.float64) - np.array([self.centering[0] * w / self.scaling, self.centering[1] * h / self.scaling])

    def pix(self, length: float) -> int:
        return int(length * self.scaling)

    def vec2pix(self, p):
        return self.pix(p[0] - self.origin[0]), self.pix(p[1] - self.origin[1])

    def fill_polygon(self, pts, color) -> None:
        """Convex polygon fill (even-odd test on the bounding box), points in pixels."""
        pts = np.asarray(pts, np.float64)
        h, w = self.img.shape[:2]
        x0, x1 = int(max(np.floor(pts[:, 0].min()), 0)), int(min(np.ceil(pts[:, 0].max()), w - 1))
        y0, y1 = int(max(np.floor(pts[:, 1].min()), 0)), int(min(np.ceil(pts[:, 1].max()), h - 1))
        if x0 > x1 or y0 > y1:
            return
        xs, ys = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
        inside = np.ones(xs.shape, bool)
        nxt = np.roll(pts, -1, axis=0)
        sign = 1.0 if np.sum(pts[:, 0] * nxt[:, 1] - nxt[:, 0] * pts[:, 1]) >= 0 else -1.0  # orientation (signed area)
        for a, b in zip(pts, nxt):
            cross = (b[0] - a[0]) * (ys - a[1]) - (b[1] - a[1]) * (xs - a[0])
            inside &= cross * sign >= -1e-9
        self.img[y0:y1 + 1, x0:x1 + 1][inside] = color

    def line(self, p, q, color, width: int) -> None:
        """``pygame.draw.line`` with a width: a rectangle around the segment."""
        p, q = np.asarray(p, np.float64), np.asarray(q, np.float64)
        d = q - p
        n = np.linalg.norm(d)
        if n < 1e-9:
            return
        t = np.array([-d[1], d[0]]) / n * (max(width, 1) / 2.0)
        self.fill_polygon([p + t, q + t, q - t, p - t], color)


def _draw_lane(lane, surf: Surface) -> None:
    """``LaneGraphics.display`` (road/graphics.py:127-257)."""
    h, w = surf.img.shape[:2]
    stripes_count = int(2 * (h + w) / (STRIPE_SPACING * surf.scaling))
    s_origin, _ = lane.local_coordinates(surf.origin)
    s0 = (int(s_origin) // STRIPE_SPACING - stripes_count // 2) * STRIPE_SPACING
    for side in range(2):
        kind = lane.line_types[side]
        if kind == LineType.NONE:
            continue
        if kind == LineType.CONTINUOUS_LINE:
            starts, ends = np.array([s0]), np.array([s0 + stripes_count * STRIPE_SPACING + STRIPE_LENGTH])
        else:
            starts = s0 + np.arange(stripes_count) * STRIPE_SPACING
            ends = starts + (STRIPE_LENGTH if kind == LineType.STRIPED else STRIPE_SPACING)
        lat = (side - 0.5) * lane.width
        starts, ends = np.clip(starts, 0, lane.length), np.clip(ends, 0, lane.length)
        for a, b in zip(starts, ends):
            if abs(a - b) > 0.5 * STRIPE_LENGTH:
                surf.line(surf.vec2pix(lane.position(a, lat)), surf.vec2pix(lane.position(b, lat)), WHITE, max(surf.pix(STRIPE_WIDTH), 1))


def _lighten(color, ratio=0.68):
    return tuple(min(int(c / ratio), 255) for c in color)


def _draw_vehicle(surf: Surface, x, y, heading, color) -> None:
    """``VehicleGraphics.display`` (vehicle/graphics.py:28-140): body, two headlights, dark outline; rotation about the centre."""
    hd = heading if abs(heading) > 2 * np.pi / 180 else 0.0
    c, s = np.cos(hd), np.sin(hd)
    centre = np.array(surf.vec2pix((x, y)), np.float64)

    def rect(x0, y0, lx, ly):  # body-frame rectangle [m] -> pixel polygon
        pts = np.array([[x0, y0], [x0 + lx, y0], [x0 + lx, y0 + ly], [x0, y0 + ly]])
        return centre + (pts @ np.array([[c, s], [-s, c]])) * surf.scaling

    body = rect(-VEH_LENGTH / 2, -VEH_WIDTH / 2, VEH_LENGTH, VEH_WIDTH)
    surf.fill_polygon(rect(-VEH_LENGTH / 2 - 0.15, -VEH_WIDTH / 2 - 0.15, VEH_LENGTH + 0.3, VEH_WIDTH + 0.3), BLACK)  # 1-px outline
    surf.fill_polygon(body, color)
    surf.fill_polygon(rect(VEH_LENGTH / 2 - 0.72, -(1.4 * VEH_WIDTH) / 3, 0.72, 0.6), _lighten(color))
    surf.fill_polygon(rect(VEH_LENGTH / 2 - 0.72, (0.6 * VEH_WIDTH) / 5, 0.72, 0.6), _lighten(color))


def render_rgb(net: RoadNetwork, state, config: dict, env_index: int = 0, linear_traffic: bool = False) -> np.ndarray:
    """One frame of env ``env_index`` of a :class:`SimState`: uint8 ``[screen_height, screen_width, 3]`` like
    ``EnvViewer.get_image`` (graphics.py:168-181; pygame's (w, h) surface array is transposed there)."""
    width, height = int(config["screen_width"]), int(config["screen_height"])
    surf = Surface(width, height, float(config.get("scaling", 5.5)), config.get("centering_position", [0.5, 0.5]))
    e = env_index
    n, ego = int(state.env_i[abi.EI_NVEH, e]), int(state.env_i[abi.EI_EGO, e])
    xs, ys, hs = state.veh_d[abi.D_X, e], state.veh_d[abi.D_Y, e], state.veh_d[abi.D_HEADING, e]
    flags = state.veh_i[abi.I_FLAGS, e]
    surf.move_display_window_to((xs[ego], ys[ego]) if n else (0.0, 0.0))   # EnvViewer.window_position
    for lane in net.lanes_list():                                           # RoadGraphics.display
        _draw_lane(lane, surf)
    for s in range(n):                                                      # RoadGraphics.display_traffic (list order)
        if flags[s] & abi.FL_CRASHED:                                       # VehicleGraphics.get_color
            color = RED
        elif flags[s] & abi.FL_MDP:
            color = GREEN
        else:
            color = YELLOW if linear_traffic else BLUE
        px = surf.vec2pix((xs[s], ys[s]))
        if -50 < px[0] < width + 50 and -50 < px[1] < height + 50:         # WorldSurface.is_visible
            _draw_vehicle(surf, xs[s], ys[s], hs[s], color)
    return surf.img
