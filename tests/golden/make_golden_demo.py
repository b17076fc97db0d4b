#!/usr/bin/env python
"""Golden vectors for what surrounds MeshNet in the reference's demo (BASELINE config 1), produced by the
UNMODIFIED reference functions in the build container:

    python tests/golden/make_golden_demo.py      ->  tests/golden/demo_pipeline.npz

  joint_img        demo/run.py:150-158 applied to demo/h36m_joint_input.npy (get_bbox, process_bbox,
                   j2d_processing of the reference; cfg.MODEL.input_shape = (384, 288))
  pose3d           posenet.get_model(17, 4096, 2, 0.5) under torch.manual_seed(123) with randomised BatchNorm
                   statistics (oracle.meshnet_oracle.randomize_bn_, seed 11), eval mode, on joint_img and on 7 seeded
                   N(0,1) poses (B = 8)
  pose_combine     pose2mesh_net.py:18-19: cat(pose2d, pose3d / 1000)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import meshnet_oracle as mo  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    ref_shim.load("human36")
    import aug_utils  # noqa: E402  (reference modules)
    import coord_utils  # noqa: E402
    from models import posenet  # noqa: E402

    joint_input = np.load(os.path.join(ref_shim.REF_ROOT, "demo", "h36m_joint_input.npy"))
    shape = ref_shim._Cfg.MODEL.input_shape
    bbox = coord_utils.get_bbox(joint_input)
    bbox2 = coord_utils.process_bbox(bbox.copy())
    joint_img, _ = aug_utils.j2d_processing(joint_input.copy(), (shape[1], shape[0]), bbox2, 0, 0, None)
    joint_img = joint_img[:, :2]
    joint_img /= np.array([[shape[1], shape[0]]])
    mean, std = np.mean(joint_img, axis=0), np.std(joint_img, axis=0)
    joint_img = (joint_img.copy() - mean) / std
    joint_img = torch.Tensor(joint_img[None, :, :])                       # run.py:159 (without .cuda())

    torch.manual_seed(123)
    net = posenet.get_model(17, hid_dim=4096, num_layer=2, p_dropout=0.5, pretrained=False)
    sd = mo.randomize_bn_({("bn." + k): v for k, v in net.state_dict().items() if "batch_norm" in k}, seed=11)
    net.load_state_dict({k[3:]: v for k, v in sd.items()}, strict=False)
    net.eval()
    g = torch.Generator().manual_seed(3)
    pose2d = torch.cat([joint_img, torch.randn(7, 17, 2, generator=g)])
    with torch.no_grad():
        pose3d = net(pose2d.view(len(pose2d), -1))
    combine = torch.cat((pose2d, pose3d.reshape(-1, 17, 3) / 1000), dim=2)   # pose2mesh_net.py:18-19
    out = os.path.join(HERE, "demo_pipeline.npz")
    np.savez_compressed(out, joint_input=joint_input, bbox2=np.asarray(bbox2, dtype=np.float64),
                        joint_img=joint_img.numpy(), pose2d=pose2d.numpy(), pose3d=pose3d.numpy(),
                        pose_combine=combine.numpy())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
