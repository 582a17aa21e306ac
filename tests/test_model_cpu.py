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
