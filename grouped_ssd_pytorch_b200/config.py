"""Prior-box configurations of the reference (ssd_liverdet/data/config.py:19-157), same dict layout
(`feature_maps, min_dim, steps, min_sizes, max_sizes, aspect_ratios, variance, clip, name`) so that
`PriorBox(cfg)` accepts either these or the reference's own dicts.

`variance` is [0.1, 0.2] and `clip` True in every configuration (config.py:36,59,82,105,129,152);
MultiBoxLoss and Detect read `v2['variance']` whatever prior set the model used
(multibox_loss.py:5,44; detection_pytorch_ver_1point5.py:10,42).
"""

VARIANCE = [0.1, 0.2]


def _cfg(name, min_dim, feature_maps, steps, sizes, aspect_ratios):
    """`sizes` holds len(feature_maps)+1 anchors scales: min_sizes = sizes[:-1], max_sizes = sizes[1:]."""
    return {
        "feature_maps": list(feature_maps),
        "min_dim": min_dim,
        "steps": list(steps),
        "min_sizes": list(sizes[:-1]),
        "max_sizes": list(sizes[1:]),
        "aspect_ratios": [list(a) for a in aspect_ratios],
        "variance": list(VARIANCE),
        "clip": True,
        "name": name,
    }


_MAPS_300 = (38, 19, 10, 5, 3, 1)
_STEPS_300 = (8, 16, 32, 64, 100, 300)
_SIZES_300 = (30, 60, 111, 162, 213, 264, 315)
_MAPS_512 = (64, 32, 16, 8, 4, 2, 1)
_STEPS_512 = (8, 16, 32, 64, 128, 256, 512)
_SIZES_512 = (20, 51, 133, 215, 296, 378, 460, 542)

# SSD300, P = 8732 (config.py:114-134) — the configuration GSSD is built with (models/...group.py:48)
v2 = _cfg("v2", 300, _MAPS_300, _STEPS_300, _SIZES_300, [[2], [2, 3], [2, 3], [2, 3], [2], [2]])
# SSD300 with an extra ar=3 on conv4_3, square anchors, P = 11620 (config.py:19-41)
v2_custom = _cfg("v2_custom", 300, _MAPS_300, _STEPS_300, _SIZES_300,
                 [[2, 3], [2, 3], [2, 3], [2, 3], [2], [2]])
# square anchors only, P = 8732 (config.py:43-64)
v2_custom_squareonly = _cfg("v2_custom_squareonly", 300, _MAPS_300, _STEPS_300, _SIZES_300,
                            [[2], [2, 3], [2, 3], [2, 3], [2], [2]])
# SSD512, P = 24564 (config.py:91-110)
v2_512 = _cfg("v2_512", 512, _MAPS_512, _STEPS_512, _SIZES_512,
              [[2], [2, 3], [2, 3], [2, 3], [2, 3], [2], [2]])
# SSD512 square anchors, P = 32756 (config.py:68-87)
v2_custom_512 = _cfg("v2_custom_512", 512, _MAPS_512, _STEPS_512, _SIZES_512,
                     [[2, 3], [2, 3], [2, 3], [2, 3], [2, 3], [2], [2]])
# legacy corner-form priors, P = 7308 (config.py:137-157)
v1 = {
    "feature_maps": list(_MAPS_300),
    "min_dim": 300,
    "steps": list(_STEPS_300),
    "min_sizes": [30, 60, 114, 168, 222, 276],
    "max_sizes": [-1, 114, 168, 222, 276, 330],
    "aspect_ratios": [[1, 1, 2, 1 / 2]] + [[1, 1, 2, 1 / 2, 3, 1 / 3] for _ in range(5)],
    "variance": list(VARIANCE),
    "clip": True,
    "name": "v1",
}

ALL = {c["name"]: c for c in (v2, v2_custom, v2_custom_squareonly, v2_512, v2_custom_512, v1)}
