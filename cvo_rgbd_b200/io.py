"""File front door of the hot path: the data formats either side of align().

  read_pcd_ascii   the reference's sample clouds data/rgbd_dataset/freiburg1_desk/pcd_ds/*.pcd
                   (`FIELDS x y z rgb`, `SIZE 8 8 8 4`, `DATA ascii`, colour packed into a float's bits)
  cloud_features   the N x 5 feature rows the two frontends expect (src/pcd_generator.cpp:336-381): cvo = raw B, G, R
                   + raw gradient; acvo = H/180, S/255, V/255 + gradient * 2 / 255.  A PCD file carries no image gradient:
                   the two gradient features are 0.
  read_assoc       TUM `assoc.txt` as the reference's drivers read it (src/cvo_main.cpp:75-101)
  PoseWriter       `name tx ty tz qx qy qz qw` of accum_transform, one line per frame once `init` is set -- the identity line of frame 0 included (src/cvo_main.cpp:58-65)
  read_trajectory  TUM trajectory files (groundtruth.txt, cvo_poses_qt.txt) -> {stamp: 4x4}

Host-side I/O only; no arithmetic of the registration path lives here.
"""
import numpy as np


def read_pcd_ascii(path):
    """Returns (xyz [N,3] float32, rgb [N,3] uint8 in R,G,B order).  Only what the reference's sample files use is
    supported: ASCII data, fields x y z [rgb]; NaN rows (invalid depth) are dropped like pcl's removeNaN would."""
    fields, n_points, data_at = None, None, None
    with open(path, "r") as f:
        lines = f.read().split("\n")
    for i, line in enumerate(lines):
        t = line.strip().split()
        if not t or t[0].startswith("#"):
            continue
        key = t[0].upper()
        if key == "FIELDS":
            fields = [x.lower() for x in t[1:]]
        elif key == "POINTS":
            n_points = int(t[1])
        elif key == "DATA":
            if t[1].lower() != "ascii":
                raise ValueError("%s: only DATA ascii is supported (got %s)" % (path, t[1]))
            data_at = i + 1
            break
    if fields is None or data_at is None or fields[:3] != ["x", "y", "z"]:
        raise ValueError("%s: not a PCD file with FIELDS x y z ..." % path)
    rows = [ln.split() for ln in lines[data_at:] if ln.strip()]
    if n_points is not None:
        rows = rows[:n_points]
    xyz = np.array([[float(v) for v in r[:3]] for r in rows], dtype=np.float64).reshape(-1, 3)
    rgb = np.zeros((len(rows), 3), np.uint8)
    if "rgb" in fields:
        k = fields.index("rgb")
        packed = np.array([np.float32(r[k]) for r in rows], dtype=np.float32)
        bits = packed.view(np.uint32)  # 0x00RRGGBB in the float's bit pattern (pcl convention)
        rgb[:, 0], rgb[:, 1], rgb[:, 2] = (bits >> 16) & 255, (bits >> 8) & 255, bits & 255
    keep = np.isfinite(xyz).all(axis=1)
    return xyz[keep].astype(np.float32), rgb[keep]


def write_pcd_ascii(path, xyz, rgb):
    """Writes the same dialect read_pcd_ascii reads (used by the tests and the examples)."""
    xyz, rgb = np.asarray(xyz, np.float64), np.asarray(rgb, np.uint32)
    bits = ((rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]).astype(np.uint32)
    packed = bits.view(np.float32)
    with open(path, "w") as f:
        f.write("# .PCD v.7 - Point Cloud Data file format\nVERSION .7\nFIELDS x y z rgb\nSIZE 8 8 8 4\n"
                "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA ascii\n"
                % (len(xyz), len(xyz)))
        for p, c in zip(xyz, packed):
            f.write("%.15f %.15f %.15f %.9g\n" % (p[0], p[1], p[2], float(c)))


def _hsv_u8(c0, c1, c2):
    """8-bit HSV of cv::cvtColor(..., COLOR_RGB2HSV) with (c0, c1, c2) taken as (R, G, B): H in [0,180), S, V in [0,255]."""
    r, g, b = c0.astype(np.float64), c1.astype(np.float64), c2.astype(np.float64)
    v = np.maximum(np.maximum(r, g), b)
    mn = np.minimum(np.minimum(r, g), b)
    d = v - mn
    s = np.where(v > 0, 255.0 * d / np.maximum(v, 1e-30), 0.0)
    dd = np.maximum(d, 1e-30)
    h = np.where(v == r, 60.0 * (g - b) / dd, np.where(v == g, 120.0 + 60.0 * (b - r) / dd, 240.0 + 60.0 * (r - g) / dd))
    h = np.where(d == 0, 0.0, h)
    h = np.where(h < 0, h + 360.0, h) / 2.0
    return np.rint(h) % 180, np.rint(s), v


def cloud_features(rgb, kind="cvo", grad=None):
    """N x 5 float32 feature rows.  kind 'cvo': feature_type 1 (src/pcd_generator.cpp:359-381): B, G, R, dI_x, dI_y raw.
    kind 'acvo': feature_type 0 (:336-358): H/180, S/255, V/255, 2 dI_x / 255, 2 dI_y / 255, where the reference converts
    its BGR image with COLOR_RGB2HSV (:389), i.e. hue is computed with the red and blue channels swapped -- reproduced."""
    rgb = np.asarray(rgb)
    n = len(rgb)
    g = np.zeros((n, 2), np.float64) if grad is None else np.asarray(grad, np.float64).reshape(n, 2)
    out = np.zeros((n, 5), np.float64)
    r_, g_, b_ = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    if kind == "cvo":
        out[:, 0], out[:, 1], out[:, 2] = b_, g_, r_
        out[:, 3:] = g
    elif kind == "acvo":
        h, s, v = _hsv_u8(b_, g_, r_)  # the BGR triple read as RGB
        out[:, 0], out[:, 1], out[:, 2] = h / 180.0, s / 255.0, v / 255.0
        out[:, 3:] = g / 255.0 * 2
    else:
        raise ValueError(kind)
    return out.astype(np.float32)


def read_assoc(path):
    """[(rgb_name, rgb_path, depth_name, depth_path)] -- one tuple per non-empty line, whitespace separated, exactly
    the four tokens load_file_name() consumes (src/cvo_main.cpp:75-101)."""
    out = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            t = (t + ["", "", "", ""])[:4]
            out.append(tuple(t))
    return out


def rotation_to_quaternion(R):
    """(qx, qy, qz, qw) of a rotation matrix, Eigen::Quaternionf(Matrix3f) branch structure (Shepperd)."""
    R = np.asarray(R, np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        x, y, z = (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q = [0.0, 0.0, 0.0]
        q[i] = 0.5 * s
        s = 0.5 / s
        w = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
        x, y, z = q
    return np.array([x, y, z, w])


def quaternion_to_rotation(q):
    x, y, z, w = np.asarray(q, np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class PoseWriter:
    """cvo_poses_qt.txt / acvo_poses_qt.txt: `name tx ty tz qx qy qz qw` of accum_transform, one line per frame once the frontend's `init` is set (frame 0: identity; src/cvo_main.cpp:58-65)."""

    def __init__(self, path):
        self._f = open(path, "w")

    def write(self, name, accum_transform):
        T = np.asarray(accum_transform, np.float64)
        q = rotation_to_quaternion(T[:3, :3])
        self._f.write("%s %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n" % (name, T[0, 3], T[1, 3], T[2, 3], q[0], q[1], q[2], q[3]))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_trajectory(path):
    """{timestamp (float): 4x4 float64} from a TUM trajectory file (`stamp tx ty tz qx qy qz qw`, `#` comments)."""
    out = {}
    with open(path) as f:
        for line in f:
            line = line.replace(",", " ").strip()
            if not line or line.startswith("#"):
                continue
            t = line.split()
            if len(t) < 8:
                continue
            v = [float(x) for x in t[:8]]
            if not np.all(np.isfinite(v)):
                continue
            T = np.eye(4)
            T[:3, :3] = quaternion_to_rotation(v[4:8])
            T[:3, 3] = v[1:4]
            out[v[0]] = T
    return out
