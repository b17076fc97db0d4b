"""pose2mesh_release_b200 — the MeshNet hot path of hongsukchoi/Pose2Mesh_RELEASE, B200-native.

Public surface (mirrors the reference's modules for this path):

    meshnet.Pose2Mesh / meshnet.get_model            <- lib/models/meshnet.py
    cheby_graph_conv.graph_conv_cheby                <- lib/models/backbones/cheby_graph_conv.py
    graph.build_coarse_graphs (+ coarsening helpers) <- lib/graph_utils.py, lib/coarsening.py
    install.install()                                 rebinding overlay for an unmodified reference checkout
    dist.DataParallelStep                             one-process-per-GPU data parallel, single NCCL all-reduce

All device work is in libp2m_b200.so (csrc/, C ABI in include/p2m_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .graph import build_coarse_graphs  # noqa: F401
from .meshnet import Pose2Mesh, get_model  # noqa: F401
from .cheby_graph_conv import graph_conv_cheby  # noqa: F401

__version__ = "0.1.0"
