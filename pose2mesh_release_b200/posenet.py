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


def weight_init(module):
    """lib/models/posenet.py:6-8 (kaiming-normal Linear weights; not applied by the reference's constructor either)."""
    if isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight)


def _register(owner: nn.Module, spec):
    """Create the sub-modules of `spec` (name -> factory) in order: registration order fixes both the state_dict key
    order and the order in which the default initialisers draw from the global RNG (identical to the reference's)."""
    for name, factory in spec:
        setattr(owner, name, factory())


class Linear(nn.Module):
    """One residual stage (lib/models/posenet.py:13-39): x + w2(drop(relu(bn2(w1(drop(relu(bn1(x)))))))."""

    def __init__(self, linear_size, p_dropout=0.5):
        super().__init__()
        h = self.l_size = linear_size
        _register(self, (("relu", lambda: nn.ReLU(inplace=True)), ("dropout", lambda: nn.Dropout(p_dropout)),
                         ("w1", lambda: nn.Linear(h, h)), ("batch_norm1", lambda: nn.BatchNorm1d(h)),
                         ("w2", lambda: nn.Linear(h, h)), ("batch_norm2", lambda: nn.BatchNorm1d(h))))

    def forward(self, x):
        y = x
        for bn, lin in ((self.batch_norm1, self.w1), (self.batch_norm2, self.w2)):
            y = lin(self.dropout(self.relu(bn(y))))
        return x + y


class LinearModel(nn.Module):
    """lib/models/posenet.py:41-87: Linear(2J, H) -> `num_stage` residual stages -> Linear(H, 3J)."""

    def __init__(self, num_joint, linear_size=4096, num_stage=2, p_dropout=0.5, pretrained=False):
        super().__init__()
        self.num_joint, self.linear_size, self.num_stage, self.p_dropout = num_joint, linear_size, num_stage, p_dropout
        self.input_size, self.output_size = 2 * num_joint, 3 * num_joint          # 2-D joints in, 3-D joints out
        h = linear_size
        _register(self, (("w1", lambda: nn.Linear(self.input_size, h)),
                         ("batch_norm1", lambda: nn.BatchNorm1d(h)),               # constructed, never applied (reference :63 vs :74-87)
                         ("linear_stages", lambda: nn.ModuleList(Linear(h, p_dropout) for _ in range(num_stage))),
                         ("w2", lambda: nn.Linear(h, self.output_size)),
                         ("relu", lambda: nn.ReLU(inplace=True)), ("dropout", lambda: nn.Dropout(p_dropout))))
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
        for stage in self.linear_stages:
            y = stage(y)
        return self.w2(y)


def get_model(num_joint, hid_dim, num_layer, p_dropout, pretrained=False):
    """lib/models/posenet.py:89-92."""
    return LinearModel(num_joint, hid_dim, num_layer, p_dropout, pretrained)
