"""GPU: device-side image ingest (b2_image_resize_dev) is bit-identical to its oracle restatement and feeds detection."""
import numpy as np
import pytest
import torch

from gtsfm_b200 import synthetic as syn
from gtsfm_b200.pipeline import DeviceFrontEnd
from oracle import images_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,res", [((480, 640, 3), 360), ((1000, 1504, 3), 760), ((333, 517), 100), ((1135, 2000, 3), 760)])
def test_resize_equals_oracle(b200_ctx, shape, res):
    fe = DeviceFrontEnd(syn.superpoint_state_dict(0), ctx=b200_ctx)
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    nh, nw = images_ref.downsampled_size(shape[0], shape[1], res)
    out = fe.ingest(torch.from_numpy(img).cuda(), res).cpu().numpy()
    assert out.shape[:2] == (nh, nw)
    assert np.array_equal(out, images_ref.resize_cubic_u8(img, nh, nw))


def test_ingest_feeds_detection(b200_ctx):
    """a large synthetic frame: ingest on the device then detect == host-side oracle resize then detect."""
    fe = DeviceFrontEnd(syn.superpoint_state_dict(0), max_keypoints=2000, ctx=b200_ctx)
    big = syn.synthetic_frame(5, 960, 1280)
    dev = fe.ingest(torch.from_numpy(big).cuda(), 480)
    host = images_ref.resize_cubic_u8(big, 480, 640)
    a, b = fe.detect(dev), fe.detect(torch.from_numpy(host).cuda())
    assert len(a) == len(b) > 500 and torch.equal(a.kp, b.kp) and torch.equal(a.desc, b.desc)
