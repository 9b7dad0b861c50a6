"""Camera model feeding ``detect_rectangle`` (hover.py:157-222).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two paths:
  * ``render_rgba``  -- a CPU rasteriser of the red target box seen by the
    PyFlyt FPV camera [RECALL: PyFlyt ``Camera.view_mat`` + pybullet
    ``computeProjectionMatrixFOV``], producing the 128x128 RGBA image the
    reference's real ``detect_rectangle`` consumes.  Used when the reference's
    own hover.py is run in the loop (tests/golden/make_golden.py).
  * ``analytic_features`` -- the pin-hole projection of the four corners of
    the panel's front face, with the pixel-lattice corrections that
    ``cv2.findContours`` / ``contourArea`` / ``boundingRect`` imply.  This is
    the float64 statement of what the CUDA kernel computes per env.

Panel geometry follows ``add_reference_object`` hover.py:116-155: box half
extents (0.02, 1, 1) with visual offset (+1, 0, 0) on a body at (6, 0, 6) yawed
by atan2(-0, -6) = -pi  =>  front face x = 4.98, y in [-1, 1], z in [5, 7].
"""
from __future__ import annotations

import numpy as np

from .quadx_model import QuadXParams, quat_to_euler, quat_to_mat

# hover.py:118-147
PANEL_BODY_POS = np.array([6.0, 0.0, 6.0])
PANEL_HALF_EXTENTS = np.array([0.02, 1.0, 1.0])
PANEL_VISUAL_OFFSET = np.array([1.0, 0.0, 0.0])
PANEL_YAW = float(np.arctan2(-0.0, -6.0))


def panel_box_corners() -> np.ndarray:
    """World coordinates of the 8 corners of the red box, [8,3]."""
    c, s = np.cos(PANEL_YAW), np.sin(PANEL_YAW)
    Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    signs = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
    local = PANEL_VISUAL_OFFSET + signs * PANEL_HALF_EXTENTS
    return PANEL_BODY_POS + local @ Rz.T


def panel_front_face() -> np.ndarray:
    """The 4 corners of the face turned to the origin, [4,3], ordered
    (y-, z-), (y+, z-), (y+, z+), (y-, z+)."""
    x = 4.98
    return np.array([[x, -1.0, 5.0], [x, 1.0, 5.0], [x, 1.0, 7.0], [x, -1.0, 7.0]])


def camera_frame(pos: np.ndarray, quat: np.ndarray, p: QuadXParams):
    """PyFlyt FPV camera [RECALL]: eye at the camera link; orientation = body
    Euler angles with the pitch offset by the tilt; looks along +x of that
    frame with +z up.  Returns eye[N,3], fwd[N,3], right[N,3], up[N,3]."""
    R = quat_to_mat(quat)
    eye = pos + np.einsum("nij,j->ni", R, np.asarray(p.cam_offset, float))
    e = quat_to_euler(quat)
    roll, pitch, yaw = e[:, 0], e[:, 1] - np.deg2rad(p.cam_tilt_up_deg), e[:, 2]
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    # columns of Rz(yaw) Ry(pitch) Rx(roll)
    fwd = np.stack([cy * cp, sy * cp, -sp], axis=1)
    left = np.stack([cy * sp * sr - sy * cr, sy * sp * sr + cy * cr, cp * sr], axis=1)
    up = np.stack([cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr], axis=1)
    return eye, fwd, -left, up


def project(points: np.ndarray, eye, fwd, right, up, p: QuadXParams):
    """World points [M,3] -> continuous pixel coords px,py [N,M] (origin top-left,
    x right, y down, image spans [0,res]) and depth [N,M]."""
    d = points[None, :, :] - eye[:, None, :]
    depth = np.einsum("nmj,nj->nm", d, fwd)
    xr = np.einsum("nmj,nj->nm", d, right)
    yu = np.einsum("nmj,nj->nm", d, up)
    t = np.tan(np.deg2rad(p.cam_fov_deg) / 2.0)
    half = p.cam_res / 2.0
    safe = np.where(depth > 1e-9, depth, 1e-9)
    px = (xr / (safe * t) + 1.0) * half
    py = (1.0 - yu / (safe * t)) * half
    return px, py, depth


# ----------------------------------------------------------------------------
# rasterised path (single env)
# ----------------------------------------------------------------------------


def render_rgba(pos: np.ndarray, quat: np.ndarray, p: QuadXParams, box_corners: np.ndarray | None = None) -> np.ndarray:
    """128x128x4 uint8 image of the red box over a non-red background.  A pixel
    is red when its centre lies inside the convex silhouette of the box.
    Shading only scales the red channel (TinyRenderer multiplies the diffuse
    colour by the light terms), so G = B = 0 and R > 100 on the whole box."""
    import cv2

    corners = panel_box_corners() if box_corners is None else box_corners
    eye, fwd, right, up = camera_frame(pos[None], quat[None], p)
    px, py, depth = project(corners, eye, fwd, right, up, p)
    res = p.cam_res
    img = np.empty((res, res, 4), np.uint8)
    img[..., 0], img[..., 1], img[..., 2], img[..., 3] = 170, 190, 230, 255
    if (depth[0] <= p.cam_near).any():
        # box crosses the near plane: clip conservatively by dropping the frame
        # when entirely behind, else paint what is in front (never happens inside
        # the 3 m flight dome, the panel is >= 2 m ahead in x).
        if (depth[0] <= p.cam_near).all():
            return img
    pts = np.stack([px[0], py[0]], axis=1).astype(np.float32)
    hull = cv2.convexHull(pts).reshape(-1, 2).astype(np.float64)
    # half-plane test at pixel centres
    ys, xs = np.mgrid[0:res, 0:res]
    cx, cy = xs + 0.5, ys + 0.5
    inside_pos = np.ones((res, res), bool)
    inside_neg = np.ones((res, res), bool)
    m = hull.shape[0]
    for i in range(m):
        x0, y0 = hull[i]
        x1, y1 = hull[(i + 1) % m]
        cr = (x1 - x0) * (cy - y0) - (y1 - y0) * (cx - x0)
        inside_pos &= cr >= 0
        inside_neg &= cr <= 0
    red = inside_pos | inside_neg
    img[red, 0], img[red, 1], img[red, 2] = 200, 0, 0
    return img


# ----------------------------------------------------------------------------
# analytic path (batched) -- what the CUDA kernel implements in fp32
# ----------------------------------------------------------------------------

# A quad counts as seen only if every corner is at least this far (pixels)
# inside the image: red on the outermost pixel ring rejects the frame
# (hover.py:180-185), and pixel 0 is red when its centre 0.5 is covered.
VIS_MARGIN_PX = 0.5


def analytic_features(pos: np.ndarray, quat: np.ndarray, p: QuadXParams):
    """-> visible bool[N], centre[N,2], area[N], ratio[N] following the
    definitions of hover.py:193-213 on the pixel lattice:

      centre  mean of the 4 projected corners, moved by -0.5 px because contour
              vertices are indices of boundary *pixels* (hover.py:197-203)
      area    cv2.contourArea runs through boundary pixel centres: the polygon
              shrunk by half a pixel all round, A - P/2 + 1   (hover.py:206)
      ratio   cv2.boundingRect counts covered pixel columns / rows (hover.py:209-213)
    """
    eye, fwd, right, up = camera_frame(pos, quat, p)
    px, py, depth = project(panel_front_face(), eye, fwd, right, up, p)
    res = p.cam_res
    lo, hi = VIS_MARGIN_PX, res - VIS_MARGIN_PX
    visible = (depth > p.cam_near).all(axis=1)
    visible &= ((px >= lo) & (px <= hi) & (py >= lo) & (py <= hi)).all(axis=1)
    half = res / 2.0
    cx = (px.mean(axis=1) - 0.5) / half - 1.0
    cy = (py.mean(axis=1) - 0.5) / half - 1.0
    # shoelace area + perimeter
    xn, yn = np.roll(px, -1, axis=1), np.roll(py, -1, axis=1)
    area_px = 0.5 * np.abs((px * yn - xn * py).sum(axis=1))
    perim = np.sqrt((xn - px) ** 2 + (yn - py) ** 2).sum(axis=1)
    area = np.maximum(area_px - 0.5 * perim + 1.0, 0.0) / (res * res)
    # covered pixel columns: centres i+0.5 in [min,max]
    w = np.floor(px.max(axis=1) - 0.5) - np.ceil(px.min(axis=1) - 0.5) + 1.0
    hgt = np.floor(py.max(axis=1) - 0.5) - np.ceil(py.min(axis=1) - 0.5) + 1.0
    visible &= (w >= 2.0) & (hgt >= 2.0)
    ratio = np.where(hgt > 0, w / np.where(hgt > 0, hgt, 1.0), 0.0)
    z = np.zeros_like(cx)
    centre = np.stack([np.where(visible, cx, z), np.where(visible, cy, z)], axis=1)
    return visible, centre, np.where(visible, area, z), np.where(visible, ratio, z)


# ----------------------------------------------------------------------------
# raster path without images -- what the CUDA kernel's vision_mode = 1 implements
# ----------------------------------------------------------------------------

# the 12 edges of the box as pairs of indices into panel_box_corners() (corners differing in exactly one sign)
BOX_EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1]


def row_spans(px: np.ndarray, py: np.ndarray, res: int):
    """Pixel rows covered by the convex silhouette of projected points (px, py) [M]: for every row r whose centre line
    y = r + 0.5 crosses the silhouette, the first and last pixel column whose centre lies inside.  The silhouette's x
    interval at that y is the min / max over the box edges that cross the line.  Returns r0 and the lists (left, right);
    rows with left > right are empty."""
    ylo, yhi = float(py.min()), float(py.max())
    r0, r1 = int(np.ceil(ylo - 0.5)), int(np.floor(yhi - 0.5))
    left, right = [], []
    for r in range(r0, r1 + 1):
        y = r + 0.5
        xs = []
        for a, b in BOX_EDGES:
            y0, y1 = py[a], py[b]
            if (y0 - y) * (y1 - y) <= 0.0 and y0 != y1:
                xs.append(px[a] + (y - y0) * (px[b] - px[a]) / (y1 - y0))
        if not xs:
            left.append(1); right.append(0)
            continue
        left.append(int(np.ceil(min(xs) - 0.5))); right.append(int(np.floor(max(xs) - 0.5)))
    return r0, left, right


def raster_features_one(pos: np.ndarray, quat: np.ndarray, p: QuadXParams):
    """(visible, centre[2], area, ratio) of one pose as hover.py:157-222 reports them for the rendered frame, computed from
    the row spans of the box silhouette instead of an image:
      red_at_edges   a covered pixel in row / column 0 or res-1 rejects the frame                     hover.py:180-189
      contourArea    the contour runs through the centres of the border pixels (Suzuki: blob pixels with a background
                     4-neighbour); by Pick's theorem its area is N - B/2 - 1 (N blob pixels, B border pixels) hover.py:206
      boundingRect   covered columns x covered rows                                                   hover.py:209-213
      centre         mean of the four approxPolyDP corners -- stood in for by the mean of the four projected front-face
                     corners - 0.5 px (within 0.8 px of Douglas-Peucker, tests/test_oracle_golden.py)     hover.py:197-203
    """
    eye, fwd, right, up = camera_frame(pos[None], quat[None], p)
    px, py, depth = project(panel_box_corners(), eye, fwd, right, up, p)
    px, py, depth = px[0], py[0], depth[0]
    res = p.cam_res
    none = (False, np.zeros(2), 0.0, 0.0)
    if (depth <= p.cam_near).any():
        return none
    r0, left, right_ = row_spans(px, py, res)
    rows = [(r0 + k, l, r) for k, (l, r) in enumerate(zip(left, right_)) if l <= r]
    if not rows:
        return none
    if rows[0][0] <= 0 or rows[-1][0] >= res - 1 or min(l for _, l, _ in rows) <= 0 or max(r for _, _, r in rows) >= res - 1:
        return none  # red at the image edge (or beyond it)
    span = {r: (l, rr) for r, l, rr in rows}
    n_pix = sum(rr - l + 1 for _, l, rr in rows)
    interior = 0
    for r, l, rr in rows:
        if (r - 1) in span and (r + 1) in span:
            lo = max(l + 1, span[r - 1][0], span[r + 1][0])
            hi = min(rr - 1, span[r - 1][1], span[r + 1][1])
            interior += max(0, hi - lo + 1)
    border = n_pix - interior
    area = (n_pix - border / 2.0 - 1.0) / (res * res)
    w = max(rr for _, _, rr in rows) - min(l for _, l, _ in rows) + 1
    h = rows[-1][0] - rows[0][0] + 1
    if w < 2 or h < 2:
        return none
    # centre: the mean of the four projected front-face corners, moved by -0.5 px (contour vertices are pixel indices) -- within
    # 0.8 px of the mean of the approxPolyDP corners (measured over 1000 poses; snapping the corners to blob pixels is worse)
    fx, fy, _ = project(panel_front_face(), eye, fwd, right, up, p)
    cx, cy = float(fx.mean()) - 0.5, float(fy.mean()) - 0.5
    half = res / 2.0
    return True, np.array([cx / half - 1.0, cy / half - 1.0]), area, w / h


def raster_features(pos: np.ndarray, quat: np.ndarray, p: QuadXParams):
    """Batched wrapper of raster_features_one -> visible[N], centre[N,2], area[N], ratio[N]."""
    n = pos.shape[0]
    v, c, a, r = np.zeros(n, bool), np.zeros((n, 2)), np.zeros(n), np.zeros(n)
    for i in range(n):
        v[i], c[i], a[i], r[i] = raster_features_one(pos[i], quat[i], p)
    return v, c, a, r
