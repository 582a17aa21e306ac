"""End-to-end parity of the model forward with the tcgen05 source blocks swapped in (`gssd_forward`) against the
outputs of the unmodified reference GSSD (tests/golden/gssd_model.npz): loc / conf of all 8732 priors within the
north-star tolerance for the bf16 conv block, 1e-2 relative to the tensor's scale, and the test-phase Detect output."""
import numpy as np
import pytest
import torch

import cases
import gssd_standin as G

pytestmark = pytest.mark.gpu


def build(phase):
    from grouped_ssd_pytorch_b200 import config
    from grouped_ssd_pytorch_b200.layers import PriorBox
    g = cases.golden("gssd_model")
    seed_w, seed_x = [int(v) for v in g["seeds"]]
    net = G.StandInSSD(phase, 2, True, PriorBox(config.v2).forward())
    net.load_state_dict(G.seeded_state(net.state_dict(), seed_w))
    net.eval().cuda()
    return net, G.seeded_input(seed_x, 1).cuda(), g


def rel(a, ref):
    return float(np.abs(a - ref).max() / np.abs(ref).max())


def test_gssd_forward_matches_the_reference_model():
    from grouped_ssd_pytorch_b200.layers.modules.source_block import gssd_forward
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net, x, g = build('train')
    loc, conf, priors = gssd_forward(net, x)
    assert loc.shape == (1, 8732, 4) and conf.shape == (1, 8732, 2) and priors.shape == (8732, 4)
    e_loc, e_conf = rel(loc.cpu().numpy(), g["loc"]), rel(conf.cpu().numpy(), g["conf"])
    assert e_loc <= 1e-2 and e_conf <= 1e-2, (e_loc, e_conf)
    # and the torch forward of the same modules on the GPU (fp32 cuDNN) agrees with the reference far more tightly
    with torch.no_grad():
        l2, c2 = G.forward_torch(net, x)
    assert rel(l2.cpu().numpy(), g["loc"]) <= 1e-3 and rel(c2.cpu().numpy(), g["conf"]) <= 1e-3


def test_gssd_forward_batch_and_test_phase():
    from grouped_ssd_pytorch_b200.layers import Detect
    from grouped_ssd_pytorch_b200.layers.modules.source_block import gssd_forward
    import types
    net, x, g = build('test')
    xb = torch.cat([x, x.flip(-1)], 0)                           # batch of 2: the second image is mirrored
    net.forward = types.MethodType(gssd_forward, net)            # the drop-in form INTEGRATION.md shows
    out = net(xb)
    assert out.shape == (2, 2, 200, 5)
    net.phase = 'train'
    loc, conf, priors = net(xb)
    assert rel(loc[:1].cpu().numpy(), g["loc"]) <= 1e-2          # image 0 is unaffected by its batch neighbour
    ref = Detect.apply(2, 0, 200, 0.01, 0.45, loc, torch.softmax(conf, -1), priors.cuda())
    assert torch.equal(out, ref)
