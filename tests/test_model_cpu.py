"""CPU check that the stand-in GSSD model the GPU tests use (tests/gssd_standin.py) IS the reference's architecture:
its plain-torch forward reproduces the outputs the unmodified reference produced for the same seeds
(tests/golden/gssd_model.npz, written by tests/golden/make_golden_model.py)."""
import numpy as np
import torch

import cases
import gssd_standin as G


def test_standin_model_reproduces_the_reference_outputs():
    g = cases.golden("gssd_model")
    seed_w, seed_x = [int(v) for v in g["seeds"]]
    net = G.StandInSSD('train', 2, True, torch.from_numpy(cases.priors("v2")))
    net.load_state_dict(G.seeded_state(net.state_dict(), seed_w))
    net.eval()
    torch.set_num_threads(8)
    with torch.no_grad():
        loc, conf = G.forward_torch(net, G.seeded_input(seed_x, 1))
    assert sum(p.numel() for p in net.parameters()) == 8340084
    np.testing.assert_allclose(loc.numpy(), g["loc"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(conf.numpy(), g["conf"], rtol=1e-4, atol=1e-4)


def test_dgrad_weight_turns_the_data_gradient_into_a_convolution():
    """host-side helper of the planned conv backward: checked against autograd in float64"""
    import torch.nn.functional as F
    from grouped_ssd_pytorch_b200.layers.modules.source_block import dgrad_weight
    g0 = torch.Generator().manual_seed(5)
    for cin, cout, groups, k in [(8, 12, 4, 3), (6, 6, 1, 1), (16, 8, 2, 3)]:
        x = torch.randn(2, cin, 7, 5, dtype=torch.float64, generator=g0).requires_grad_()
        w = torch.randn(cout, cin // groups, k, k, dtype=torch.float64, generator=g0)
        y = F.conv2d(x, w, padding=k // 2, groups=groups)
        dy = torch.randn(y.shape, dtype=torch.float64, generator=g0)
        (gx,) = torch.autograd.grad(y, x, dy)
        wt = dgrad_weight(w, groups)
        assert wt.shape == (cin, cout // groups, k, k)
        np.testing.assert_allclose(F.conv2d(dy, wt, padding=k // 2, groups=groups).numpy(), gx.numpy(), rtol=0, atol=1e-12)
