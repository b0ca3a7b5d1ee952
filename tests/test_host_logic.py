"""CPU checks of the host-side mirror modules a caller imports instead of the reference's (no GPU, no kernels)."""
import numpy as np

from util import golden


def test_product_transforms_match_reference_golden():
    """velocity_b200.transforms against the reference's own outputs (utils/transforms.py:7-73): rpy2dcm, dcm2rpy,
    transform and the 3-element pseudo-quaternion helpers quat2dcm / dcm2quat."""
    from velocity_b200 import transforms as T

    g = golden("transforms")
    for r, c, b in zip(g["rpy"], g["dcm"], g["back"]):
        assert np.allclose(T.rpy2dcm(r), c, rtol=0, atol=1e-15)
        assert np.allclose(T.dcm2rpy(c), b, rtol=0, atol=1e-15)
    for r, out in zip(g["rpy"][:4], g["xf"]):
        assert np.allclose(T.transform(g["X"], r, g["t"]), out, rtol=0, atol=1e-13)
    for c, q, c2 in zip(g["dcm"], g["quat"], g["quat_dcm"]):
        assert np.allclose(T.dcm2quat(c), q, rtol=0, atol=1e-13)
        assert np.allclose(T.quat2dcm(q), c2, rtol=0, atol=1e-13)


def test_product_projection_helpers_match_reference_golden():
    """velocity_b200.NLS.fzK / fzC and common.world2image / pixel2uvec against the reference's outputs."""
    from velocity_b200 import common as Cm

    g = golden("projection")
    K, a, R = g["K"], g["a"], g["R"]
    assert np.allclose(Cm.pscale(a @ K), g["fzK"], rtol=0, atol=1e-12)                                     # fzK, utils/NLS.py:71-78
    cam = np.concatenate([R, np.array([[0.1, 0.2, 0.3]])]) @ K
    assert np.allclose(Cm.pscale(Cm.addcol1(a) @ cam), g["fzC"], rtol=0, atol=1e-12)                      # fzC, utils/NLS.py:80-86
    assert np.allclose(Cm.world2image(K, R, np.array([0.1, 0.2, 0.3]), a), g["world2image"], rtol=0, atol=1e-12)
    assert np.allclose(Cm.pixel2uvec(K, g["fzK"]), g["pixel2uvec"], rtol=0, atol=1e-13)
