"""The native host-buffer pipeline (gssd_pipe_*, grouped_ssd_pytorch_b200/pipeline.py) must return exactly what the
public `layers` API returns for the same inputs — it launches the same kernels — with several steps in flight."""
import numpy as np
import pytest
import torch

import cases
from grouped_ssd_pytorch_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_pipeline_equals_the_layers_api_with_steps_in_flight():
    from grouped_ssd_pytorch_b200.layers import Detect, MultiBoxLoss
    from grouped_ssd_pytorch_b200.pipeline import HostPipeline
    pri = torch.from_numpy(cases.priors("v2")).cuda()
    P, B = pri.shape[0], 4
    pipe = HostPipeline(B, pri, num_classes=2, top_k=200, depth=3, conf_thresh=0.2, nms_thresh=0.45)
    r = syn.rng(77)
    steps = []
    for i in range(7):
        tg = syn.targets(r, B, 1, 6)
        steps.append(dict(loc=syn.loc(r, B, P), conf=syn.conf_logits(r, B, P, 2), scores=syn.detect_scores(r, B, P, 2, -4.0),
                          targets=[torch.from_numpy(t) for t in tg]))
    bufs = [pipe.host_buffers() for _ in range(3)]
    crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
    tickets = []
    for i, s in enumerate(steps):
        b = bufs[i % 3]
        if i >= 3:
            pipe.wait(tickets[i - 3])                            # the buffers of step i-3 are free again
            check(steps[i - 3], b, pipe, tickets[i - 3], crit, pri, Detect, grads=False)
        b.loc.copy_(torch.from_numpy(s["loc"])); b.conf.copy_(torch.from_numpy(s["conf"])); b.scores.copy_(torch.from_numpy(s["scores"]))
        tickets.append(pipe.submit(b, s["targets"]))
    for i in range(len(steps) - 3, len(steps)):
        pipe.wait(tickets[i])
        check(steps[i], bufs[i % 3], pipe, tickets[i], crit, pri, Detect, grads=True)
    pipe.close()


def check(s, b, pipe, ticket, crit, pri, Detect, grads):
    loc = torch.from_numpy(s["loc"]).cuda().requires_grad_()
    conf = torch.from_numpy(s["conf"]).cuda().requires_grad_()
    ll, lc = crit((loc, conf, pri), s["targets"])
    (ll + lc).backward()
    assert float(b.losses[0]) == float(ll) and float(b.losses[1]) == float(lc)
    out = Detect.apply(2, 0, 200, 0.2, 0.45, loc.detach(), torch.from_numpy(s["scores"]).cuda(), pri)
    assert torch.equal(b.detections, out.cpu())
    if grads:                                                    # still resident: no later step has reused the slot
        gl, gc = pipe.grads(ticket)
        assert torch.equal(gl, loc.grad) and torch.equal(gc, conf.grad)


def test_pipeline_errors_mirror_the_reference():
    from grouped_ssd_pytorch_b200.pipeline import HostPipeline
    pri = torch.from_numpy(cases.priors("small")).cuda()
    with pytest.raises(ValueError):
        HostPipeline(2, pri, nms_thresh=0.0)
    pipe = HostPipeline(2, pri, depth=2)
    b = pipe.host_buffers()
    with pytest.raises(IndexError):
        pipe.submit(b, [torch.zeros(0, 5), torch.zeros(1, 5)])
    pipe.close()
