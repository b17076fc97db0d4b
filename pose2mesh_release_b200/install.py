"""Drop the B200 MeshNet into a running copy of the reference.

    import __init_path                      # the reference's sys.path hack (main/__init_path.py)
    import pose2mesh_release_b200.install as p2m; p2m.install()
    import core.base                        # Trainer / Tester now build the B200 MeshNet

``install()`` rebinds ``models.meshnet.Pose2Mesh`` / ``get_model`` and
``models.backbones.cheby_graph_conv.graph_conv_cheby`` and swaps ``graph_utils.build_coarse_graphs``
for the native-matching builder; ``uninstall()`` restores the originals.  Nothing in the reference
tree is modified on disk.
"""
from __future__ import annotations

import importlib

_saved = {}


def install(replace_graph_builder: bool = True):
    from . import cheby_graph_conv as my_conv
    from . import graph as my_graph
    from . import meshnet as my_meshnet

    ref_meshnet = importlib.import_module("models.meshnet")
    ref_conv = importlib.import_module("models.backbones.cheby_graph_conv")
    _saved.setdefault("meshnet", (ref_meshnet.Pose2Mesh, ref_meshnet.get_model, ref_meshnet.graph_conv_cheby))
    _saved.setdefault("conv", ref_conv.graph_conv_cheby)
    ref_meshnet.Pose2Mesh = my_meshnet.Pose2Mesh
    ref_meshnet.get_model = my_meshnet.get_model
    ref_meshnet.graph_conv_cheby = my_conv.graph_conv_cheby
    ref_conv.graph_conv_cheby = my_conv.graph_conv_cheby
    if replace_graph_builder:
        ref_gu = importlib.import_module("graph_utils")
        _saved.setdefault("graph", ref_gu.build_coarse_graphs)
        ref_gu.build_coarse_graphs = my_graph.build_coarse_graphs


def uninstall():
    if "meshnet" in _saved:
        ref_meshnet = importlib.import_module("models.meshnet")
        ref_meshnet.Pose2Mesh, ref_meshnet.get_model, ref_meshnet.graph_conv_cheby = _saved.pop("meshnet")
    if "conv" in _saved:
        importlib.import_module("models.backbones.cheby_graph_conv").graph_conv_cheby = _saved.pop("conv")
    if "graph" in _saved:
        importlib.import_module("graph_utils").build_coarse_graphs = _saved.pop("graph")
