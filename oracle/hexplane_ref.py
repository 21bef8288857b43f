"""ORACLE — test infrastructure only.  CPU/PyTorch restatement of the reference's
`deform_network.forward` (scene/deformation.py:252-253 -> forward_dynamic2 :158-199, quat2mat
:417-438; scene/hexplane.py:19-108 normalize_aabb / grid_sample_wrapper / interpolate_ms_features;
utils/graphics_utils.py:117-140 batch_quaternion_multiply) for the Stereo-Blur configuration.
Pinned: tests/golden/hexplane_w128.npz holds inputs / state_dict / outputs of the reference module
itself (tests/golden/make_golden.py); tests/test_oracle.py::test_hexplane_ref_matches_reference_golden."""
import itertools
import math

import torch
import torch.nn.functional as F


def hexplane_features(pts, times, aabb, grids):
    """pts [N,3], times [N,1], aabb [2,3], grids[l][p] [1,C,H,W] -> [N, C*L]"""
    p = torch.clamp((pts - aabb[0]) * (2.0 / (aabb[1] - aabb[0])) - 1.0, -1.0, 1.0)
    p = torch.cat([p, times], dim=-1)
    feats = []
    for level in grids:
        prod = 1.0
        for plane, comb in zip(level, itertools.combinations(range(4), 2)):
            coords = p[:, list(comb)].view(1, 1, -1, 2)
            s = F.grid_sample(plane, coords, align_corners=True, mode="bilinear", padding_mode="border")
            prod = prod * s.view(plane.shape[1], -1).t()
        feats.append(prod)
    return torch.cat(feats, dim=-1)


def quat2mat5(q4):
    nq = torch.cat([torch.ones_like(q4[:, :1]), q4], dim=1)
    nq = nq / nq.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def quat_mul_normalised(q1, q2):
    w = q1[:, 0] * q2[:, 0] - q1[:, 1] * q2[:, 1] - q1[:, 2] * q2[:, 2] - q1[:, 3] * q2[:, 3]
    x = q1[:, 0] * q2[:, 1] + q1[:, 1] * q2[:, 0] + q1[:, 2] * q2[:, 3] - q1[:, 3] * q2[:, 2]
    y = q1[:, 0] * q2[:, 2] - q1[:, 1] * q2[:, 3] + q1[:, 2] * q2[:, 0] + q1[:, 3] * q2[:, 1]
    z = q1[:, 0] * q2[:, 3] + q1[:, 1] * q2[:, 2] - q1[:, 2] * q2[:, 1] + q1[:, 3] * q2[:, 0]
    q = torch.stack((w, x, y, z), dim=1)
    return q / q.norm(dim=1, keepdim=True)


def deform_forward_ref(net, point, scales, rotations, times_sel):
    """net: an object with the reference's attribute tree (deformation_net.grid.grids/aabb,
    .feature_out, .pos_deform, .scales_deform, .rotations_deform)."""
    d = net.deformation_net
    feat = hexplane_features(point[:, :3], times_sel[:, :1], d.grid.aabb, d.grid.grids)
    hidden = d.feature_out(feat)
    dx = d.pos_deform(hidden)
    pts = point[:, :3] + dx[:, 0:3]
    pts = quat2mat5(dx[:, 3:]).bmm(pts.unsqueeze(-1)).squeeze(-1)
    ds = torch.clamp(d.scales_deform(hidden), -math.log(100), math.log(100))
    new_scales = scales[:, :3] + ds
    dr = d.rotations_deform(hidden)
    new_rots = quat_mul_normalised(rotations[:, :4] + dr, dx[:, 3:])
    return pts, new_scales, new_rots
