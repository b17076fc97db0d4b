"""PoseNet, the 2-D -> 3-D pose lifter in front of MeshNet: drop-in for the reference's ``models.posenet``
(lib/models/posenet.py:13-87) — SURVEY.md §8 row f1.

``LinearModel`` keeps the reference's constructor, attribute names and ``state_dict`` keys (``w1``, ``batch_norm1``
(constructed but unused by the reference's forward), ``linear_stages.<i>.{w1,batch_norm1,w2,batch_norm2}``, ``w2``)
so PoseNet checkpoints load unchanged.  In eval mode (what FlatPose2Mesh uses for inference, demo/run.py:168) the
forward runs in libp2m_b200.so (``p2m_posenet_forward``: fp32 GEMMs with the BatchNorm / ReLU / residual fused into
their epilogues).  In training mode (dropout + batch statistics, lib/core/base.py PoseNet pre-training) the module
falls through to the same torch ops the reference runs — on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


def weight_init(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight)


class Linear(nn.Module):
    """One residual stage (lib/models/posenet.py:13-39)."""

    def __init__(self, linear_size, p_dropout=0.5):
        super().__init__()
        self.l_size = linear_size
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(p_dropout)
        self.w1 = nn.Linear(self.l_size, self.l_size)
        self.batch_norm1 = nn.BatchNorm1d(self.l_size)
        self.w2 = nn.Linear(self.l_size, self.l_size)
        self.batch_norm2 = nn.BatchNorm1d(self.l_size)

    def forward(self, x):
        y = self.w1(self.dropout(self.relu(self.batch_norm1(x))))
        y = self.w2(self.dropout(self.relu(self.batch_norm2(y))))
        return x + y


class LinearModel(nn.Module):
    """lib/models/posenet.py:41-87."""

    def __init__(self, num_joint, linear_size=4096, num_stage=2, p_dropout=0.5, pretrained=False):
        super().__init__()
        self.linear_size = linear_size
        self.p_dropout = p_dropout
        self.num_stage = num_stage
        self.num_joint = num_joint
        self.input_size = num_joint * 2
        self.output_size = num_joint * 3
        self.w1 = nn.Linear(self.input_size, self.linear_size)
        self.batch_norm1 = nn.BatchNorm1d(self.linear_size)   # constructed, never applied (reference :63, :74-87)
        self.linear_stages = nn.ModuleList([Linear(self.linear_size, self.p_dropout) for _ in range(num_stage)])
        self.w2 = nn.Linear(self.linear_size, self.output_size)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(self.p_dropout)
        if pretrained:
            raise RuntimeError("pretrained PoseNet weights are loaded by the reference's own checkpoint code "
                               "(funcs_utils.load_checkpoint); load_state_dict() them into this module")

    def _native_params(self):
        stages = (_lib.PoseNetStage * self.num_stage)()
        keep = []
        for i, st in enumerate(self.linear_stages):
            vals = [st.w1.weight, st.w1.bias, st.w2.weight, st.w2.bias,
                    st.batch_norm1.weight, st.batch_norm1.bias, st.batch_norm1.running_mean, st.batch_norm1.running_var,
                    st.batch_norm2.weight, st.batch_norm2.bias, st.batch_norm2.running_mean, st.batch_norm2.running_var]
            for (name, _), t in zip(_lib.PoseNetStage._fields_, vals):
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError("PoseNet parameters must be contiguous float32")
                setattr(stages[i], name, t.data_ptr())
            keep.append(vals)
        p = _lib.PoseNetParams()
        p.num_joint, p.hidden, p.num_stage = self.num_joint, self.linear_size, self.num_stage
        p.w1_w, p.w1_b = self.w1.weight.data_ptr(), self.w1.bias.data_ptr()
        p.w2_w, p.w2_b = self.w2.weight.data_ptr(), self.w2.bias.data_ptr()
        p.stages = stages
        p._keep = (stages, keep)
        return p

    def forward_native(self, x: torch.Tensor, with_combine: bool = False):
        """Eval forward in libp2m_b200.so.  x [B, 2J] (or [B, J, 2]) on the module's CUDA device -> pose3d [B, 3J];
        with_combine additionally returns pose_combine = cat(pose2d, pose3d / 1000) [B, J, 5]
        (lib/models/pose2mesh_net.py:18-19)."""
        lib = _lib.load()
        if not x.is_cuda:
            raise RuntimeError("pose2mesh_release_b200 runs on CUDA (sm_100a) only; got a CPU tensor")
        x = x.reshape(len(x), -1).contiguous().float()
        if x.shape[1] != self.input_size:
            raise ValueError(f"PoseNet expects {self.input_size} inputs per pose, got {x.shape[1]}")
        dev, B = x.device, x.shape[0]
        if self.w1.weight.device != dev:
            raise RuntimeError("parameters and input live on different devices")
        out = torch.empty((B, self.output_size), device=dev, dtype=torch.float32)
        comb = torch.empty((B, self.num_joint, 5), device=dev, dtype=torch.float32) if with_combine else None
        nbytes = lib.p2m_posenet_workspace_bytes(B, self.linear_size)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        params = self._native_params()
        with torch.cuda.device(dev):
            _lib.check(lib.p2m_posenet_forward(C.byref(params), x.data_ptr(), out.data_ptr(),
                                               None if comb is None else comb.data_ptr(), B, ws.data_ptr(), nbytes,
                                               torch.cuda.current_stream(dev).cuda_stream), "p2m_posenet_forward")
        return (out, comb) if with_combine else out

    def forward(self, x):
        if not self.training and x.is_cuda and not (torch.is_grad_enabled() and x.requires_grad):
            return self.forward_native(x)
        y = self.w1(x)                     # training (dropout, batch statistics): the reference's own op sequence
        for i in range(self.num_stage):
            y = self.linear_stages[i](y)
        return self.w2(y)


def get_model(num_joint, hid_dim, num_layer, p_dropout, pretrained=False):
    """lib/models/posenet.py:89-92."""
    return LinearModel(num_joint, hid_dim, num_layer, p_dropout, pretrained)
