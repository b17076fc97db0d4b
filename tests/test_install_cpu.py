"""install() / uninstall(): the rebinding overlay over an UNMODIFIED reference checkout.
CPU-only; skipped where /root/reference does not exist (the GPU box)."""
import importlib
import sys

import pytest
import torch

from helpers import graph_from_fixture
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture()
def reference():
    gu, coarsening, meshnet, conv = ref_shim.load("human36")
    import pose2mesh_release_b200.install as inst

    yield gu, meshnet, conv, inst
    inst.uninstall()


def test_install_rebinds_every_holder_and_uninstall_restores(reference):
    gu, ref_meshnet, ref_conv, inst = reference
    import pose2mesh_release_b200 as p2m
    from pose2mesh_release_b200 import cheby_graph_conv as my_conv
    from pose2mesh_release_b200 import graph as my_graph

    orig = (ref_meshnet.Pose2Mesh, ref_meshnet.get_model, ref_conv.graph_conv_cheby, gu.build_coarse_graphs)
    # a module imported BEFORE install() that holds its own binding, like the reference's datasets and demo/run.py
    # (`from graph_utils import build_coarse_graphs`, data/Human36M/dataset.py:13, demo/run.py:18)
    holder = type(sys)("early_holder")
    holder.build_coarse_graphs = gu.build_coarse_graphs
    holder.graph_conv_cheby = ref_conv.graph_conv_cheby
    sys.modules["early_holder"] = holder
    try:
        inst.install()
        assert ref_meshnet.Pose2Mesh is p2m.Pose2Mesh and ref_meshnet.get_model is p2m.get_model
        assert ref_conv.graph_conv_cheby is my_conv.graph_conv_cheby
        assert ref_meshnet.graph_conv_cheby is my_conv.graph_conv_cheby     # meshnet.py:9 imported the name
        assert gu.build_coarse_graphs is my_graph.build_coarse_graphs
        assert holder.build_coarse_graphs is my_graph.build_coarse_graphs
        assert holder.graph_conv_cheby is my_conv.graph_conv_cheby
        inst.install()                                                       # idempotent
        inst.uninstall()
        assert (ref_meshnet.Pose2Mesh, ref_meshnet.get_model, ref_conv.graph_conv_cheby, gu.build_coarse_graphs) == orig
        assert holder.build_coarse_graphs is orig[3] and holder.graph_conv_cheby is orig[2]
    finally:
        sys.modules.pop("early_holder", None)


def test_flat_pose2mesh_builds_the_b200_meshnet(reference):
    """The reference's own wrapper (lib/models/pose2mesh_net.py:5,14 -> meshnet.get_model) picks the overlay up,
    with the reference's state_dict keys under the `pose2mesh.` prefix."""
    gu, ref_meshnet, ref_conv, inst = reference
    import pose2mesh_release_b200 as p2m

    mats, _ = graph_from_fixture("smpl_small")
    torch.manual_seed(123)
    with ref_shim.cpu_cuda_noop():
        ref_model = ref_meshnet.get_model(5, 3, [m.copy() for m in mats])    # the reference mutates its list
    ref_keys = sorted(ref_model.state_dict().keys())
    inst.install()
    pose2mesh_net = importlib.import_module("models.pose2mesh_net")
    flat = pose2mesh_net.get_model(17, [m.copy() for m in mats])
    assert isinstance(flat.pose2mesh, p2m.Pose2Mesh)
    from pose2mesh_release_b200 import posenet as my_posenet

    assert isinstance(flat.pose_lifter, my_posenet.LinearModel)               # the front half is rebound as well
    got = sorted(k[len("pose2mesh."):] for k in flat.state_dict() if k.startswith("pose2mesh."))
    assert got == ref_keys
    flat.pose2mesh.load_state_dict(ref_model.state_dict())                    # reference checkpoints load unchanged
    # drop-in users get the tensor-core path without touching the module (P2M_PRECISION overrides)
    from pose2mesh_release_b200 import _lib

    assert flat.pose2mesh._hier.precision == _lib.default_precision() == _lib.P2M_PREC_FP16X3_TC
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            flat.pose2mesh(torch.zeros(1, 17, 5))


def test_precision_env_override(monkeypatch):
    from pose2mesh_release_b200 import _lib

    monkeypatch.setenv("P2M_PRECISION", "fp32")
    assert _lib.default_precision() == _lib.P2M_PREC_FP32_SIMT
    monkeypatch.setenv("P2M_PRECISION", "bf16")
    with pytest.raises(RuntimeError, match="P2M_PRECISION"):
        _lib.default_precision()


def test_posenet_mirror_has_the_reference_state_dict(reference):
    """models.posenet.LinearModel (lib/models/posenet.py:41-72) and the B200 mirror: same keys, shapes and — under the
    same seed — the same initial values, so PoseNet checkpoints load unchanged."""
    import importlib

    from pose2mesh_release_b200 import pose2mesh_net as my_flat
    from pose2mesh_release_b200 import posenet as my_posenet

    ref_posenet = importlib.import_module("models.posenet")
    torch.manual_seed(5)
    ref = ref_posenet.get_model(17, hid_dim=64, num_layer=2, p_dropout=0.5, pretrained=False)
    torch.manual_seed(5)
    mine = my_posenet.get_model(17, hid_dim=64, num_layer=2, p_dropout=0.5, pretrained=False)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), k
    mats, _ = graph_from_fixture("smpl_small")
    flat = my_flat.get_model(17, [m.copy() for m in mats])
    assert {k.split(".")[0] for k in flat.state_dict()} == {"pose_lifter", "pose2mesh"}   # pose2mesh_net.py:13-14
