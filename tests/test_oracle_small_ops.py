"""CPU: the oracle's camera_rays and the product's tensor-op ray set-up against the REAL reference's golden vectors
(tests/golden/small_ops.npz), and -- when /root/reference is present -- against the reference modules live."""
import pytest
import torch

from oracle import ref_shim, tracer as otr
from tests.util import load_golden


def test_camera_rays_against_reference_golden():
    from nefii_b200.utils import rend_util
    g = load_golden("small_ops.npz")
    uv, K = torch.from_numpy(g["cr_uv"]), torch.from_numpy(g["cr_K"])
    p44, p7 = torch.from_numpy(g["cr_pose44"]), torch.from_numpy(g["cr_pose7"])
    d, c = otr.camera_rays(uv, p44, K)
    assert torch.equal(d, torch.from_numpy(g["cr_dirs44"])) and torch.equal(c, torch.from_numpy(g["cr_cam44"]))
    for pose, tag in ((p44, "44"), (p7, "7")):
        d, c = rend_util.get_camera_params(uv, pose, K)          # CPU tensors: the tensor-op path
        assert torch.equal(d, torch.from_numpy(g["cr_dirs" + tag])), tag
        assert torch.equal(c, torch.from_numpy(g["cr_cam" + tag])), tag
    # the quaternion and the matrix form describe the same camera
    assert torch.allclose(torch.from_numpy(g["cr_dirs7"]), torch.from_numpy(g["cr_dirs44"]), atol=1e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_sample_network_golden_is_the_reference():
    ref_shim.install()
    from model.sample_network import SampleNetwork
    g = load_golden("small_ops.npz")
    inp = [torch.from_numpy(g["sn_" + k]) for k in ("s", "s0", "grad", "t0", "cam", "dirs")]
    assert torch.equal(SampleNetwork()(*inp), torch.from_numpy(g["sn_x"]))
