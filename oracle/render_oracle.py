"""CPU ORACLE (test infrastructure, NOT product code) for the analytic renderer: a numpy restatement, in float64, of what the fork's
PyBullet.render produces for the primitive scenes (reference panda_gym/pybullet.py:70-107 camera, :149-264 render + deprojection):
one ray per pixel centre against oriented boxes / z-cylinders, OpenGL depth-buffer values, the reference's inv(P V) deprojection of the
pixel-corner NDC grid, the depth < 0.99 and workspace filters.  Only tests/ may import this module.

Parity unpinned against the real engine: pybullet's camera conventions are restated from knowledge of bullet3 (b3ComputeViewMatrixFromYawPitchRoll,
b3ComputeProjectionMatrixFOV), and the reference draws the robot's visual meshes, which are not in /root/reference -- the robot here is the
set of boxes the physics uses."""
import numpy as np


def camera(target, distance, yaw, pitch, roll):
    y, p, r = np.radians([yaw, pitch, roll])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(r), 0, np.sin(r)], [0, 1, 0], [-np.sin(r), 0, np.cos(r)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]])
    R = Rz @ Ry @ Rx
    target = np.asarray(target, float)
    eye, up = R @ np.array([0.0, -distance, 0.0]) + target, R @ np.array([0.0, 0.0, 1.0])
    f = (target - eye) / np.linalg.norm(target - eye)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    return eye, f, s, np.cross(s, f)


def _hit_box(o, d, c, R, h):
    """o [3], d [P,3]; box centre c, world axes R (columns), half extents h.  Returns entry distance [P] (inf = miss)."""
    lo, ld = R.T @ (o - c), d @ R
    with np.errstate(divide="ignore", invalid="ignore"):
        a, b = (-h - lo) / ld, (h - lo) / ld
    par = ld == 0
    a = np.where(par, -np.inf, a); b = np.where(par, np.inf, b)
    miss_par = (par & (np.abs(lo) > h)).any(1)
    t0, t1 = np.minimum(a, b).max(1), np.maximum(a, b).min(1)
    return np.where((t0 <= t1) & (t0 > 0) & ~miss_par, t0, np.inf)


def _hit_cyl(o, d, c, R, r, hz):
    lo, ld = R.T @ (o - c), d @ R
    best = np.full(len(d), np.inf)
    a, b, cc = ld[:, 0] ** 2 + ld[:, 1] ** 2, lo[0] * ld[:, 0] + lo[1] * ld[:, 1], lo[0] ** 2 + lo[1] ** 2 - r * r
    disc = b * b - a * cc
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (-b - np.sqrt(np.maximum(disc, 0))) / a
        z = lo[2] + t * ld[:, 2]
        best = np.where((a > 0) & (disc >= 0) & (t > 0) & (np.abs(z) <= hz), t, best)
        sgn = np.where(ld[:, 2] > 0, -1.0, 1.0)
        tc = (sgn * hz - lo[2]) / ld[:, 2]
        x, y = lo[0] + tc * ld[:, 0], lo[1] + tc * ld[:, 1]
        best = np.where((ld[:, 2] != 0) & (tc > 0) & (x * x + y * y <= r * r) & (tc < best), tc, best)
    return best


def render(prims, width, height, target=(0, 0, 0), distance=1.4, yaw=45, pitch=-30, roll=0, crop=True, near=0.1, far=100.0):
    """prims: list of (kind 'box'|'cyl', id, centre[3], R[3,3], half[3]).  Returns depth [H,W], segmentation [H,W], points [H,W,3], valid [H,W]."""
    eye, f, s, u = camera(target, distance, yaw, pitch, roll)
    th, asp = np.tan(np.radians(30.0)), width / height
    rows, cols = np.mgrid[0:height, 0:width]
    xn, yn = (cols.reshape(-1) + 0.5) * (2.0 / width) - 1.0, 1.0 - (rows.reshape(-1) + 0.5) * (2.0 / height)
    d = f[None, :] + (xn * th * asp)[:, None] * s[None, :] + (yn * th)[:, None] * u[None, :]
    best, seg = np.full(width * height, np.inf), np.zeros(width * height, np.uint8)
    for kind, pid, c, R, h in prims:
        t = _hit_box(eye, d, np.asarray(c, float), np.asarray(R, float), np.asarray(h, float)) if kind == "box" else _hit_cyl(eye, d, np.asarray(c, float), np.asarray(R, float), h[0], h[2])
        closer = t < best
        best, seg = np.where(closer, t, best), np.where(closer, pid, seg).astype(np.uint8)
    hit = (best >= near) & (best <= far)
    with np.errstate(divide="ignore", invalid="ignore"):
        zn = np.where(hit, ((far + near) - 2 * far * near / best) / (far - near), 1.0)
    depth = np.where(hit, 0.5 * (zn + 1.0), 1.0)
    seg = np.where(hit, seg, 0).astype(np.uint8)
    # the reference's deprojection: NDC of the pixel corner grid (np.mgrid[-1:1:2/h, -1:1:2/w], y flipped), z = 2 depth - 1, through inv(P V)
    xc, yc = cols.reshape(-1) * (2.0 / width) - 1.0, -(rows.reshape(-1) * (2.0 / height) - 1.0)
    ze = np.where(hit, best, 1.0)
    pts = eye[None, :] + ze[:, None] * f[None, :] + (xc * th * asp * ze)[:, None] * s[None, :] + (yc * th * ze)[:, None] * u[None, :]
    valid = hit & (depth < 0.99)
    if crop:
        valid &= (pts[:, 2] > 0.0) & (pts[:, 2] < 0.67) & (pts[:, 0] > -0.5) & (pts[:, 0] < 0.2)
    return depth.reshape(height, width), seg.reshape(height, width), pts.reshape(height, width, 3), valid.reshape(height, width)


def deproject_reference(depth, width, height, target=(0, 0, 0), distance=1.4, yaw=45, pitch=-30, roll=0, near=0.1, far=100.0):
    """The reference's own deprojection arithmetic (pybullet.py:205-241) on a depth-buffer image: 4x4 matrices, inv(P V), homogeneous divide.
    Used to check that the closed form above (and in the kernel) IS that arithmetic."""
    eye, f, s, u = camera(target, distance, yaw, pitch, roll)
    V = np.eye(4); V[0, :3], V[1, :3], V[2, :3] = s, u, -f; V[:3, 3] = -V[:3, :3] @ eye
    ys = 1.0 / np.tan(np.radians(60.0) / 2)
    P = np.array([[ys / (width / height), 0, 0, 0], [0, ys, 0, 0], [0, 0, (near + far) / (near - far), 2 * near * far / (near - far)], [0, 0, -1, 0]])
    T = np.linalg.inv(P @ V)
    y, x = np.mgrid[-1:1:2 / height, -1:1:2 / width]
    y = y * -1.0
    pix = np.stack([x.reshape(-1), y.reshape(-1), 2 * depth.reshape(-1) - 1, np.ones(width * height)], axis=1)
    pts = (T @ pix.T).T
    return (pts / pts[:, 3:4])[:, :3].reshape(height, width, 3)
