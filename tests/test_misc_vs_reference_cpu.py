"""Host-side helpers of the path against the reference's own functions (util/misc.py:291-336, :477-480; util/box_ops.py:18-74),
imported through oracle/ref_shim.py -- skipped where the reference tree (or its staged copy baseline/_ref) is absent."""
import pytest
import torch

from oracle import ref_shim


@pytest.fixture(scope="module")
def ref():
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    return ref_shim.load_reference()


def test_nested_tensor_from_ragged_list(ref):
    from spe_b200.util import misc
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randn(3, 37, 52, generator=g), torch.randn(3, 64, 40, generator=g), torch.randn(3, 10, 64, generator=g)]
    a = misc.nested_tensor_from_tensor_list(imgs)
    b = ref.misc.nested_tensor_from_tensor_list(imgs)
    assert torch.equal(a.tensors, b.tensors) and torch.equal(a.mask, b.mask)
    ta, ma = a.decompose()
    assert ta.shape == (3, 3, 64, 64) and ma.dtype == torch.bool and bool(ma[0, 37:, :].all()) and not bool(ma[0, :37, :52].any())


def test_inverse_sigmoid(ref):
    from spe_b200.util import misc
    x = torch.tensor([0.0, 1e-7, 1e-5, 0.3, 0.5, 0.999999, 1.0, 1.2, -0.1])
    assert torch.equal(misc.inverse_sigmoid(x), ref.misc.inverse_sigmoid(x))


def test_box_conversions(ref):
    from spe_b200.util import box_ops
    g = torch.Generator().manual_seed(1)
    b = torch.rand(50, 4, generator=g)
    assert torch.equal(box_ops.box_cxcywh_to_xyxy(b), ref.box_ops.box_cxcywh_to_xyxy(b))
    xy = ref.box_ops.box_cxcywh_to_xyxy(b)
    assert torch.equal(box_ops.box_xyxy_to_cxcywh(xy), ref.box_ops.box_xyxy_to_cxcywh(xy))
