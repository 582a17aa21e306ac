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
    # the C entry point checks the host-side row offsets itself (callers other than pipeline.py)
    import ctypes as C
    from grouped_ssd_pytorch_b200 import _lib
    lib = _lib.load()
    gt = torch.tensor([[0.1, 0.1, 0.5, 0.5, 0.0]] * 3).pin_memory()
    b.loc.zero_(); b.conf.zero_(); b.scores.zero_()

    def submit(off, sum_g, g_max):
        o = torch.tensor(off, dtype=torch.int32).pin_memory()
        return lib.gssd_pipe_submit(pipe._h, b.loc.data_ptr(), b.conf.data_ptr(), b.scores.data_ptr(), gt.data_ptr(), o.data_ptr(),
                                    sum_g, g_max, b.losses.data_ptr(), b.detections.data_ptr())
    assert submit([0, 1, 2], 3, 2) == _lib.ERR_ARG            # offsets do not end at sum_g
    assert submit([0, 3, 3], 3, 3) == _lib.ERR_EMPTY          # image without ground truth
    assert submit([0, 2, 3], 3, 1) == _lib.ERR_ARG            # more rows in an image than g_max
    assert submit([1, 2, 3], 3, 2) == _lib.ERR_ARG
    t = submit([0, 2, 3], 3, 2)
    assert t >= 0
    pipe.wait(t)
    # begin without finish: the slot cannot be taken again
    st = C.c_void_p()
    o = torch.tensor([0, 2, 3], dtype=torch.int32).pin_memory()
    begin = lambda: lib.gssd_pipe_begin(pipe._h, b.loc.data_ptr(), b.conf.data_ptr(), b.scores.data_ptr(), gt.data_ptr(), o.data_ptr(),
                                        3, 2, b.detections.data_ptr(), C.byref(st))
    t1 = begin(); t2 = begin()
    assert t1 >= 0 and t2 >= 0 and begin() == _lib.ERR_ARG
    for t in (t1, t2):
        assert lib.gssd_pipe_finish(pipe._h, t, None, 0, b.losses.data_ptr()) == 0
    pipe.wait(t2)
    pipe.close()
