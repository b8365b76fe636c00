"""helpers for the frustum-cull extension tests: a clip_from_world matrix and the planes() of culling_bind_group.rs:25-38 in numpy"""
import numpy as np


def look_at_perspective(eye, target, fovy_deg=60.0, aspect=16 / 9, near=0.1, far=1e6):
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    f = target - eye
    f /= np.linalg.norm(f)
    up = np.array([0.0, 1.0, 0.0]) if abs(f[1]) < 0.99 else np.array([1.0, 0.0, 0.0])
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[:3, 3] = -view[:3, :3] @ eye
    t = 1.0 / np.tan(np.radians(fovy_deg) / 2)
    proj = np.zeros((4, 4))
    proj[0, 0], proj[1, 1] = t / aspect, t
    proj[2, 2], proj[2, 3] = far / (near - far), near * far / (near - far)
    proj[3, 2] = -1.0
    return (proj @ view).astype(np.float32)       # row-major numpy matrix M with clip = M @ world


def planes_numpy(m):
    """planes(): row3 +- row0, row3 +- row1, row3 - row2 (f32)"""
    m = np.asarray(m, np.float32)
    out = np.zeros((5, 4), np.float32)
    for i in range(5):
        row = m[i // 2]
        out[i] = m[3] + row if (i & 1) == 0 and i != 4 else m[3] - row
    return out
