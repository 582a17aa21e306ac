"""Generate tests/golden/gssd_model.npz: the UNMODIFIED reference GSSD model (build_ssd, ssd_type gssd, batch_norm,
2 classes) run end to end on a seeded 4-phase input with seeded parameters, in eval mode.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_model.py       (build container only: needs /root/reference)

Stores the model outputs loc[1,8732,4] / conf[1,8732,2] (models/ssd_multiphase_custom_group.py:392-396).  Input and
parameters are regenerated in the tests from the seeds (tests/gssd_standin.py); this script also checks that the
stand-in model used on the GPU box has exactly the reference's state_dict (keys, shapes) and produces the same
outputs through plain torch.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/ssd_liverdet")
warnings.filterwarnings("ignore")

import torch  # noqa: E402

dcn = types.ModuleType("dcn_v2")
dcn._DCNv2 = type("_DCNv2", (), {"apply": staticmethod(lambda *a: None)})
sys.modules["dcn_v2"] = dcn
mpl = types.ModuleType("matplotlib")
mpl.use = lambda *a, **k: None
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")

from models.ssd_multiphase_custom_group import build_ssd  # noqa: E402  (reference)

import gssd_standin as G  # noqa: E402

SEED_W, SEED_X = 71, 72
torch.set_num_threads(8)


def main():
    net = build_ssd('train', 300, 2, True, 4, 4, 1, True, False, False, 0, 1, False, False, 1)
    sd = net.state_dict()
    stand = G.StandInSSD('train', 2, True, net.priors.detach().clone())
    sd2 = stand.state_dict()
    assert list(sd.keys()) == list(sd2.keys()), "stand-in state_dict keys differ from the reference"
    assert all(tuple(sd[k].shape) == tuple(sd2[k].shape) for k in sd), "stand-in state_dict shapes differ"
    state = G.seeded_state(sd, SEED_W)
    net.load_state_dict(state)
    net.eval()
    x = G.seeded_input(SEED_X, 1)
    with torch.no_grad():
        loc, conf, priors = net(x)
    print("loc", tuple(loc.shape), float(loc.abs().max()), "conf", tuple(conf.shape), float(conf.abs().max()))
    path = os.path.join(HERE, "gssd_model.npz")
    np.savez_compressed(path, loc=loc.numpy().astype(np.float32), conf=conf.numpy().astype(np.float32),
                        seeds=np.array([SEED_W, SEED_X]))
    print("gssd_model %8.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
