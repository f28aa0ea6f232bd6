"""Seeded synthetic RGB-D frame pairs shaped like the reference's inputs.

The reference selects ~3000 high-gradient pixels per frame (num_want,
src/pcd_generator.cpp:22) and back-projects them with the TUM fr1 intrinsics
(src/pcd_generator.cpp:251-257).  We imitate that output contract -- N x 3 xyz f32
plus N x 5 features f32 -- with a "desk-like" scene: three planes and three boxes,
points drawn along edges / texture lines, depth 0.7-2.0 m inside the fr1 frustum.
Feature flavours follow src/pcd_generator.cpp:336-381:
  "cvo"  (feature type 1): raw BGR in [0,255] + raw gradients
  "acvo" (feature type 0): HSV/[180,255,255] in [0,1] + gradient*2/255

Everything is numpy.random.default_rng(seed); seed = 1000*cfg + pair_index.
"""
import numpy as np

FX, FY, CX, CY, W, H = 517.3, 516.5, 318.6, 255.3, 640, 480
Z_MIN, Z_MAX = 0.70, 2.00


def _rect(origin, e1, e2):
    return (np.asarray(origin, float), np.asarray(e1, float), np.asarray(e2, float))


def _scene(rng):
    """Rectangles (origin, edge1, edge2) + albedo; fixed layout jittered by the seed."""
    rects = []
    # desk y = 0.35, back wall z = 1.7, side wall x = -0.7 (camera looks along +z, y down)
    rects.append(_rect([-0.7, 0.35, 0.80], [1.4, 0, 0], [0, 0, 0.90]))
    rects.append(_rect([-0.7, -0.45, 1.7], [1.4, 0, 0], [0, 0.80, 0]))
    rects.append(_rect([-0.7, -0.45, 0.80], [0, 0, 0.90], [0, 0.80, 0]))
    for _ in range(3):  # boxes standing on the desk: top, front and one side face
        sx, sy, sz = rng.uniform(0.10, 0.30), rng.uniform(0.08, 0.30), rng.uniform(0.10, 0.25)
        cx, cz = rng.uniform(-0.55, 0.55 - sx), rng.uniform(0.95, 1.6 - sz)
        y0 = 0.35 - sy
        rects.append(_rect([cx, y0, cz], [sx, 0, 0], [0, 0, sz]))          # top
        rects.append(_rect([cx, y0, cz], [sx, 0, 0], [0, sy, 0]))          # front (faces camera)
        side_x = cx if cx > 0 else cx + sx
        rects.append(_rect([side_x, y0, cz], [0, 0, sz], [0, sy, 0]))      # side
    albedo_bgr = rng.uniform(30, 225, size=(len(rects), 3))
    return rects, albedo_bgr


def _segments(rects, rng):
    """Edges of every rectangle plus random interior texture lines; returns (p0, p1, rect_id)."""
    p0, p1, rid = [], [], []
    for k, (o, e1, e2) in enumerate(rects):
        corners = [o, o + e1, o + e1 + e2, o + e2]
        for a in range(4):
            p0.append(corners[a]); p1.append(corners[(a + 1) % 4]); rid.append(k)
        n_tex = 1
        for _ in range(n_tex):
            a, b = rng.uniform(0, 1, 2), rng.uniform(0, 1, 2)
            p0.append(o + a[0] * e1 + a[1] * e2); p1.append(o + b[0] * e1 + b[1] * e2); rid.append(k)
    return np.array(p0), np.array(p1), np.array(rid)


def _bgr_to_hsv01(bgr):
    """OpenCV-style HSV (H in [0,180], S,V in [0,255]) scaled to [0,1] as src/pcd_generator.cpp:341-343."""
    b, g, r = bgr[:, 0], bgr[:, 1], bgr[:, 2]
    v = np.max(bgr, axis=1)
    mn = np.min(bgr, axis=1)
    diff = v - mn
    s = np.where(v > 0, 255.0 * diff / np.maximum(v, 1e-12), 0.0)
    h = np.zeros_like(v)
    nz = diff > 1e-12
    rmax = nz & (v == r)
    gmax = nz & (v == g) & ~rmax
    bmax = nz & ~rmax & ~gmax
    h[rmax] = 60.0 * (g[rmax] - b[rmax]) / diff[rmax]
    h[gmax] = 120.0 + 60.0 * (b[gmax] - r[gmax]) / diff[gmax]
    h[bmax] = 240.0 + 60.0 * (r[bmax] - g[bmax]) / diff[bmax]
    h = np.where(h < 0, h + 360.0, h) / 2.0
    return np.stack([h / 180.0, s / 255.0, v / 255.0], axis=1)


def _draw(n, p0, p1, rid, albedo_bgr, rng, flavour):
    seg_len = np.linalg.norm(p1 - p0, axis=1)
    prob = seg_len / seg_len.sum()
    pts = np.zeros((0, 3)); ids = np.zeros((0,), int)
    while pts.shape[0] < n:
        m = int((n - pts.shape[0]) * 1.5) + 16
        s = rng.choice(len(prob), size=m, p=prob)
        t = rng.uniform(0, 1, size=(m, 1))
        q = p0[s] + t * (p1[s] - p0[s]) + rng.normal(0, 0.003, size=(m, 3))
        z = q[:, 2]
        u = FX * q[:, 0] / z + CX
        vv = FY * q[:, 1] / z + CY
        ok = (z > Z_MIN) & (z < Z_MAX) & (u >= 0) & (u < W) & (vv >= 0) & (vv < H)
        pts = np.concatenate([pts, q[ok]]); ids = np.concatenate([ids, rid[s][ok]])
    pts, ids = pts[:n], ids[:n]
    bgr = np.clip(albedo_bgr[ids] + rng.normal(0, 6.0, size=(n, 3)), 0, 255)
    grad = rng.normal(0, 20.0, size=(n, 2))
    if flavour == "cvo":
        feat = np.concatenate([bgr, grad], axis=1)
    else:
        feat = np.concatenate([_bgr_to_hsv01(bgr), grad * 2.0 / 255.0], axis=1)
    return pts, feat


def _rotvec_to_R(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_pair(seed, n_fixed=3000, n_moving=3000, flavour="cvo", motion_scale=1.0):
    """Returns dict(x_pos, x_feat, y_pos, y_feat, T_gt) with f32 arrays.

    T_gt (4x4, f64) maps moving-frame points into the fixed frame -- the quantity the
    reference's `transform` converges to.  Motion statistics follow
    fr1_desk: rotation ~N(0,(0.0207/sqrt3 rad)^2) per axis,
    translation ~N(0,(0.0163/sqrt3 m)^2) per axis, clipped at 0.146 rad / 0.0425 m.
    """
    rng = np.random.default_rng(seed)
    rects, albedo = _scene(rng)
    p0, p1, rid = _segments(rects, rng)
    w = np.clip(rng.normal(0, 0.0207 / np.sqrt(3), 3) * motion_scale, -0.146, 0.146)
    t = np.clip(rng.normal(0, 0.0163 / np.sqrt(3), 3) * motion_scale, -0.0425, 0.0425)
    Rg = _rotvec_to_R(w)
    x_pos, x_feat = _draw(n_fixed, p0, p1, rid, albedo, rng, flavour)
    y_fix, y_feat = _draw(n_moving, p0, p1, rid, albedo, rng, flavour)
    y_pos = (y_fix - t) @ Rg  # R^T (p - t), row-vector form
    # 2 mm noise along the viewing ray of the moving camera
    ray = y_pos / np.linalg.norm(y_pos, axis=1, keepdims=True)
    y_pos = y_pos + ray * rng.normal(0, 0.002, size=(n_moving, 1))
    T_gt = np.eye(4)
    T_gt[:3, :3] = Rg
    T_gt[:3, 3] = t
    return dict(x_pos=x_pos.astype(np.float32), x_feat=x_feat.astype(np.float32),
                y_pos=y_pos.astype(np.float32), y_feat=y_feat.astype(np.float32), T_gt=T_gt)


def config_pair(cfg, pair_index=0):
    """The five BASELINE.json configs. seed = 1000*cfg + pair_index."""
    seed = 1000 * cfg + pair_index
    if cfg == 1:
        return make_pair(seed, 500, 500, "cvo")
    if cfg == 2:
        return make_pair(seed, 3000, 3000, "cvo")
    if cfg == 3:
        return make_pair(seed, 3000, 3000, "acvo")
    if cfg == 4:
        rng = np.random.default_rng(seed + 7_000_000)
        n, m = rng.integers(2700, 3301, 2)
        return make_pair(seed, int(n), int(m), "cvo")
    if cfg == 5:
        return make_pair(seed, 10000, 10000, "cvo")
    raise ValueError("cfg must be 1..5")


def make_frame(seed, w=640, h=480, texture=1.0):
    """A seeded synthetic RGB-D frame shaped like the reference's image inputs (src/cvo_main.cpp:104-107): an 8-bit
    3-channel image of random overlapping rectangles, discs and line segments (edges and texture for the DSO pixel
    selector, src/pcd_generator.cpp:122-164) with sensor noise, and a 16-bit depth image (TUM scaling 5000/m) of a
    tilted plane with boxes and a few zero-depth holes.  `texture` < 1 thins the structure out (low-texture frames
    exercise the selector's re-selection and the Canny top-up).  Returns (img3 uint8[h,w,3], depth uint16[h,w])."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    base = rng.uniform(60, 190, 3)
    img[:] = base + 25.0 * np.stack([np.sin(xx / 97.0 + i) * np.cos(yy / 71.0 - i) for i in range(3)], axis=-1)
    depth_m = 1.0 + 0.6 * (yy / h) + 0.2 * (xx / w)
    n_shapes = max(1, int(40 * texture))
    for _ in range(n_shapes):
        col = rng.uniform(10, 245, 3)
        if rng.uniform() < 0.6:
            x0, y0 = rng.integers(0, w - 20), rng.integers(0, h - 20)
            x1, y1 = x0 + rng.integers(15, 220), y0 + rng.integers(15, 160)
            m = (xx >= x0) & (xx < x1) & (yy >= y0) & (yy < y1)
            depth_m[m] = rng.uniform(0.7, 1.9)
        else:
            cx, cy, r = rng.integers(0, w), rng.integers(0, h), rng.integers(8, 70)
            m = (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
        img[m] = col + 12.0 * np.sin((xx[m] + 2 * yy[m]) / rng.uniform(3, 15))[:, None]
    for _ in range(int(25 * texture)):
        x0, y0, x1, y1 = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(0, w), rng.uniform(0, h)
        d = np.abs((y1 - y0) * xx - (x1 - x0) * yy + x1 * y0 - y1 * x0) / max(np.hypot(y1 - y0, x1 - x0), 1.0)
        img[d < rng.uniform(0.7, 2.0)] = rng.uniform(0, 255, 3)
    img += rng.normal(0, 2.0 * texture + 0.3, img.shape)
    img3 = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    depth = np.clip(np.rint(depth_m * 5000.0), 0, 65535).astype(np.uint16)
    for _ in range(6):  # holes: the sensor returns 0 where it has no reading
        x0, y0 = rng.integers(0, w - 40), rng.integers(0, h - 30)
        depth[y0:y0 + rng.integers(5, 30), x0:x0 + rng.integers(5, 40)] = 0
    return img3, depth
