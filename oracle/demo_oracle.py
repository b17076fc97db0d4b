"""TEST INFRASTRUCTURE ONLY — CPU restatement of what surrounds MeshNet in the reference's demo / model wrapper.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.

    demo/run.py:150-160 (optimize_cam_param)         2-D pose -> normalised network input
      lib/coord_utils.py:7-18,21-39,42-66            get_center_scale, get_bbox, process_bbox
      lib/aug_utils.py:51-64,140-179,182-195         j2d_processing, get_affine_transform, affine_transform
    lib/models/posenet.py:13-87                      PoseNet (LinearModel: 2 residual stages of 4096 units)
    lib/models/pose2mesh_net.py:16-22                FlatPose2Mesh.forward: cat(pose2d, pose3d / 1000) -> MeshNet
    lib/core/base.py:130-131,201-204; demo/run.py:170-171   vertex gather + joint regression

Parity status: PINNED — tests/golden/demo_pipeline.npz holds the outputs of the unmodified reference functions
(tests/golden/make_golden_demo.py) and tests/test_oracle_golden.py checks this restatement against them.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

INPUT_SHAPE = (384, 288)  # cfg.MODEL.input_shape = (height, width), lib/core/config.py:52


def get_bbox(joint_img: np.ndarray) -> np.ndarray:
    """coord_utils.py:21-39: tight box [xmin, ymin, w, h] as float32."""
    x, y = joint_img[:, 0], joint_img[:, 1]
    xmin, ymin, xmax, ymax = min(x), min(y), max(x), max(y)
    xc, w = (xmin + xmax) / 2.0, xmax - xmin
    yc, h = (ymin + ymax) / 2.0, ymax - ymin
    return np.array([xc - 0.5 * w, yc - 0.5 * h, w, h]).astype(np.float32)


def process_bbox(bbox: np.ndarray, aspect_ratio=None, scale=1.0):
    """coord_utils.py:42-66: sanitise, then grow to the network's aspect ratio (width / height)."""
    x, y, w, h = bbox
    x1, y1, x2, y2 = x, y, x + (w - 1), y + (h - 1)
    if not (w * h > 0 and x2 >= x1 and y2 >= y1):
        return None
    bbox = np.array([x1, y1, x2 - x1, y2 - y1])
    w, h = bbox[2], bbox[3]
    cx, cy = bbox[0] + w / 2.0, bbox[1] + h / 2.0
    if aspect_ratio is None:
        aspect_ratio = INPUT_SHAPE[1] / INPUT_SHAPE[0]
    if w > aspect_ratio * h:
        h = w / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    bbox[2], bbox[3] = w * scale, h * scale
    bbox[0], bbox[1] = cx - bbox[2] / 2.0, cy - bbox[3] / 2.0
    return bbox


def affine_from_bbox(bbox: np.ndarray, res) -> np.ndarray:
    """get_center_scale (coord_utils.py:7-18) + get_affine_transform with rot = 0 (aug_utils.py:140-173).
    With no rotation the three point pairs cv2.getAffineTransform solves for describe a uniform scaling by
    res[0] / box_width that maps the box centre to the patch centre; restated in closed form (float32 points,
    float64 solve, like OpenCV)."""
    x, y, w, h = bbox
    center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
    scale = np.array([w * 1.0, h * 1.0], dtype=np.float32)
    src_w, dst_w, dst_h = scale[0], res[0], res[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + np.array([0, src_w * -0.5], np.float32)
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)

    def third(a, b):  # aug_utils.py:182-185
        d = a - b
        return b + np.array([-d[1], d[0]], dtype=np.float32)

    src[2], dst[2] = third(src[0], src[1]), third(dst[0], dst[1])
    a = np.concatenate([src.astype(np.float64), np.ones((3, 1))], axis=1)      # [3,3] rows (x, y, 1)
    return np.linalg.solve(a, dst.astype(np.float64)).T                        # [2,3]


def j2d_processing(kp: np.ndarray, res, bbox) -> np.ndarray:
    """aug_utils.py:51-64 with rot = 0, no flip.  The reference transforms IN PLACE in the input's dtype: the demo
    fixture is int64 (demo/h36m_joint_input.npy), so every transformed coordinate is truncated towards zero when it
    is written back into `kp` before the final astype('float32')."""
    trans = affine_from_bbox(bbox, res)
    kp = kp.copy()
    for i in range(kp.shape[0]):
        pt = np.array([kp[i, 0], kp[i, 1], 1.0])
        kp[i, :2] = np.dot(trans, pt)[:2]
    return kp.astype("float32")


def normalize_pose2d(joint_input: np.ndarray) -> np.ndarray:
    """demo/run.py:150-158: 2-D joints in pixels -> zero-mean / unit-std coordinates of the 288 x 384 input box."""
    bbox2 = process_bbox(get_bbox(joint_input).copy())
    joint_img = j2d_processing(joint_input.copy(), (INPUT_SHAPE[1], INPUT_SHAPE[0]), bbox2)[:, :2]
    joint_img = joint_img / np.array([[INPUT_SHAPE[1], INPUT_SHAPE[0]]])
    mean, std = np.mean(joint_img, axis=0), np.std(joint_img, axis=0)
    return ((joint_img.copy() - mean) / std).astype(np.float32)


# ----------------------------------------------------------------------------------------------- PoseNet
def posenet_init_state_dict(num_joint: int, hid: int = 4096, num_stage: int = 2) -> Dict[str, torch.Tensor]:
    """State dict with the reference's names / shapes / default initialisers, drawing from the global torch RNG in
    the reference's construction order (posenet.py:41-72: w1, batch_norm1 (unused in forward), the stages' w1, bn1,
    w2, bn2, then w2)."""
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, fin, fout):
        m = torch.nn.Linear(fin, fout)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()

    def bn(name, f):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(f), torch.zeros(f)
        sd[name + ".running_mean"], sd[name + ".running_var"] = torch.zeros(f), torch.ones(f)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    lin("w1", num_joint * 2, hid)
    bn("batch_norm1", hid)
    for s in range(num_stage):
        lin(f"linear_stages.{s}.w1", hid, hid)
        bn(f"linear_stages.{s}.batch_norm1", hid)
        lin(f"linear_stages.{s}.w2", hid, hid)
        bn(f"linear_stages.{s}.batch_norm2", hid)
    lin("w2", hid, num_joint * 3)
    return sd


def posenet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_stage: int = 2, eps: float = 1e-5) -> torch.Tensor:
    """posenet.py:74-87 + :27-39 in eval mode (running-stat BatchNorm, dropout off): x [B, 2J] -> [B, 3J]."""
    y = F.linear(x, sd["w1.weight"], sd["w1.bias"])
    for s in range(num_stage):
        p = f"linear_stages.{s}."
        h = F.batch_norm(y, sd[p + "batch_norm1.running_mean"], sd[p + "batch_norm1.running_var"],
                         sd[p + "batch_norm1.weight"], sd[p + "batch_norm1.bias"], False, 0.1, eps)
        h = F.linear(F.relu(h), sd[p + "w1.weight"], sd[p + "w1.bias"])
        h = F.batch_norm(h, sd[p + "batch_norm2.running_mean"], sd[p + "batch_norm2.running_var"],
                         sd[p + "batch_norm2.weight"], sd[p + "batch_norm2.bias"], False, 0.1, eps)
        h = F.linear(F.relu(h), sd[p + "w2.weight"], sd[p + "w2.bias"])
        y = y + h
    return F.linear(y, sd["w2.weight"], sd["w2.bias"])


def flat_pose2mesh_input(pose2d: torch.Tensor, pose3d: torch.Tensor) -> torch.Tensor:
    """pose2mesh_net.py:18-19: MeshNet input [B, J, 5] = cat(pose2d, pose3d / 1000)."""
    j = pose2d.shape[1]
    return torch.cat((pose2d, pose3d.reshape(-1, j, 3) / 1000), dim=2)


def regress_joints(mesh: torch.Tensor, perm_reverse, n_vertex: int, joint_regressor: torch.Tensor):
    """base.py:130-131 / run.py:170-171: gather the real vertices, then joints = J_regressor @ vertices."""
    verts = mesh[:, torch.as_tensor(np.asarray(perm_reverse[:n_vertex])), :]
    return verts, torch.matmul(joint_regressor, verts)
