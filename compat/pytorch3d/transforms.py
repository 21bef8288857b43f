"""pytorch3d.transforms stand-ins (eval.py:17; wxyz real-first quaternions)."""
import torch


def quaternion_to_matrix(q):
    w, x, y, z = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (y * y + z * z), two_s * (x * y - z * w), two_s * (x * z + y * w),
                     two_s * (x * y + z * w), 1 - two_s * (x * x + z * z), two_s * (y * z - x * w),
                     two_s * (x * z - y * w), two_s * (y * z + x * w), 1 - two_s * (x * x + y * y)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def matrix_to_quaternion(m):
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    w = torch.sqrt(torch.clamp(1 + m00 + m11 + m22, min=1e-12)) / 2
    x = (m[..., 2, 1] - m[..., 1, 2]) / (4 * w)
    y = (m[..., 0, 2] - m[..., 2, 0]) / (4 * w)
    z = (m[..., 1, 0] - m[..., 0, 1]) / (4 * w)
    return torch.stack((w, x, y, z), -1)
