"""Seeded synthetic inputs for the multibox hot path.

numpy only (frozen `RandomState` streams), so that the golden generator, the tests, the bench and
the CPU baseline all see bit-identical inputs on any box.
"""
import numpy as np

SEED = 1111  # train_lesion_multiphase_v2.py:4-5


def rng(seed=SEED):
    return np.random.RandomState(seed)


def targets(r, batch, g_min=1, g_max=5):
    """list of [G,5] float32 rows (xmin,ymin,xmax,ymax,label=0), the data_custom_v2.py:260-263 format.

    centre ~ U[0.1,0.9]^2, size ~ U[0.03,0.28]^2, corners clamped to [0,1]."""
    out = []
    for _ in range(batch):
        g = int(r.randint(g_min, g_max + 1))
        c = r.uniform(0.1, 0.9, size=(g, 2))
        wh = r.uniform(0.03, 0.28, size=(g, 2))
        t = np.concatenate([np.clip(c - wh / 2, 0, 1), np.clip(c + wh / 2, 0, 1), np.zeros((g, 1))], 1)
        out.append(t.astype(np.float32))
    return out


def loc(r, batch, num_priors, sigma=0.5):
    return (r.standard_normal((batch, num_priors, 4)) * sigma).astype(np.float32)


def conf_logits(r, batch, num_priors, num_classes=2):
    return r.standard_normal((batch, num_priors, num_classes)).astype(np.float32)


def softmax(x):
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


def detect_scores(r, batch, num_priors, num_classes=2, shift=-4.0):
    """softmax scores; `shift` is added to the non-background logits: -4 -> ~3 % of priors above 0.2
    ("sparse-realistic"), 0 -> ~84 % ("dense stress")."""
    x = conf_logits(r, batch, num_priors, num_classes)
    x[..., 1:] += np.float32(shift)
    return softmax(x)
