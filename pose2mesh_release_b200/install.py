"""Drop the B200 MeshNet into a running copy of the reference.

    import __init_path                      # the reference's sys.path hack (main/__init_path.py)
    import pose2mesh_release_b200.install as p2m; p2m.install()
    import core.base                        # Trainer / Tester now build the B200 MeshNet

``install()`` rebinds ``models.meshnet.Pose2Mesh`` / ``get_model``,
``models.backbones.cheby_graph_conv.graph_conv_cheby`` and ``models.posenet.LinearModel`` / ``get_model`` (the
PoseNet in front of MeshNet: the reference's own ``FlatPose2Mesh`` then runs both halves natively in eval mode) and
swaps ``graph_utils.build_coarse_graphs`` for the native-matching builder; ``uninstall()`` restores the originals.  Nothing in the reference
tree is modified on disk.
"""
from __future__ import annotations

import importlib
import sys

_saved = {}
_rebound = []  # (module, attribute, original) for `from graph_utils import build_coarse_graphs`-style holders


def _rebind_holders(original, replacement):
    """Modules imported BEFORE install() that did ``from graph_utils import build_coarse_graphs`` (the
    reference's datasets and demo/run.py do) hold their own reference to the original function: patch
    those bindings too, so that the overlay is never half applied."""
    for name, mod in list(sys.modules.items()):
        if mod is None or name.startswith("pose2mesh_release_b200"):
            continue
        try:
            items = list(vars(mod).items())
        except TypeError:
            continue
        for attr, val in items:
            if val is original:
                setattr(mod, attr, replacement)
                _rebound.append((mod, attr, original))


def install(replace_graph_builder: bool = True):
    from . import cheby_graph_conv as my_conv
    from . import graph as my_graph
    from . import meshnet as my_meshnet
    from . import posenet as my_posenet

    ref_posenet = importlib.import_module("models.posenet")
    _saved.setdefault("posenet", (ref_posenet.LinearModel, ref_posenet.get_model))
    ref_posenet.LinearModel = my_posenet.LinearModel
    ref_posenet.get_model = my_posenet.get_model
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
        _rebind_holders(_saved["graph"], my_graph.build_coarse_graphs)   # includes graph_utils itself
        ref_gu.build_coarse_graphs = my_graph.build_coarse_graphs
    _rebind_holders(_saved["conv"], my_conv.graph_conv_cheby)


def uninstall():
    while _rebound:
        mod, attr, original = _rebound.pop()
        setattr(mod, attr, original)
    if "meshnet" in _saved:
        ref_meshnet = importlib.import_module("models.meshnet")
        ref_meshnet.Pose2Mesh, ref_meshnet.get_model, ref_meshnet.graph_conv_cheby = _saved.pop("meshnet")
    if "posenet" in _saved:
        ref_posenet = importlib.import_module("models.posenet")
        ref_posenet.LinearModel, ref_posenet.get_model = _saved.pop("posenet")
    if "conv" in _saved:
        importlib.import_module("models.backbones.cheby_graph_conv").graph_conv_cheby = _saved.pop("conv")
    if "graph" in _saved:
        importlib.import_module("graph_utils").build_coarse_graphs = _saved.pop("graph")
