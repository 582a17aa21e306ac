"""The evaluator on the GPU (csrc/evalap.cu) against the numpy oracle (oracle/evalap.py) and against what
the reference's own test_net returned for the same detections (tests/golden/evalap.npz): the filtered rows, the TP / FP code of
every detection at every threshold and the global score order are bit-exact; AP / IoBB agree to 1e-12 (float64 on both sides;
the kernel adds the curve's terms in a different order than np.sum)."""
import numpy as np
import pytest
import torch

import cases
from oracle import evalap as EA

pytestmark = pytest.mark.gpu
AP_LIST, IOBB_LIST = [0.3, 0.5, 0.7], [0.3, 0.5, 0.7]


def _case(tag):
    g = cases.golden("evalap")
    W, H, thresh = g[tag + "/meta"]
    gt_off = g[tag + "/gt_off"]
    gts = [g[tag + "/gt"][gt_off[i]:gt_off[i + 1]] for i in range(len(gt_off) - 1)]
    return g, g[tag + "/out"], gts, float(W), float(H), float(thresh)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_collect_and_ap_match_the_reference_evaluator(tag):
    from grouped_ssd_pytorch_b200.layers.functions import ap_iobb, collect_detections, evaluate_detections
    g, out, gts, W, H, thresh = _case(tag)
    dev_out = torch.from_numpy(out).cuda()
    want_rows, want_off = EA.collect_detections(out, W, H, thresh)
    rows = collect_detections(dev_out, W, H, thresh)
    assert rows.dtype == np.float32 and np.array_equal(rows, want_rows)
    d_rows, d_off = collect_detections(dev_out, W, H, thresh, as_numpy=False)
    assert np.array_equal(d_off.cpu().numpy(), want_off)
    for metric in (1, 0):
        o_ap, o_iobb, o_code, o_order = EA.ap_iobb(want_rows, gts, AP_LIST, IOBB_LIST, use_07_metric=bool(metric))
        ap, iobb, tp, order = ap_iobb(d_rows, d_off, gts, AP_LIST, IOBB_LIST, use_07_metric=bool(metric), details=True)
        assert np.array_equal(tp.cpu().numpy(), o_code), "TP / FP codes differ from the oracle"
        assert np.array_equal(order.cpu().numpy().astype(np.int64), o_order), "global score order differs"
        np.testing.assert_allclose(ap, g["%s/ap_%d" % (tag, metric)], rtol=1e-12, atol=1e-15)       # the reference's own numbers
        np.testing.assert_allclose(iobb, g["%s/iobb_%d" % (tag, metric)], rtol=1e-12, atol=1e-15)
    # the batched front end (several Detect output batches, one image size)
    halves = [dev_out[:len(gts) // 2], dev_out[len(gts) // 2:]]
    ap2, iobb2 = evaluate_detections(halves, (W, H), gts, thresh=thresh, ap_list=AP_LIST, iobb_list=IOBB_LIST)
    np.testing.assert_allclose(ap2, g["%s/ap_1" % tag], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(iobb2, g["%s/iobb_1" % tag], rtol=1e-12, atol=1e-15)


def test_equal_scores_keep_image_rank_order_and_edge_cases():
    """the tie contract (equal scores inside an image and across images), images without boxes, images without detections, a
    box detected twice (second one is a false positive), thresholds met with equality (strict >)"""
    from grouped_ssd_pytorch_b200.layers.functions import ap_iobb, collect_detections
    r = np.random.RandomState(4)
    I, K = 30, 200
    out = np.zeros((I, 2, K, 5), np.float32)
    gts = []
    for i in range(I):
        g = 0 if i % 7 == 3 else int(r.randint(1, 6))
        c = r.uniform(0.2, 0.8, (g, 2)); wh = r.uniform(0.1, 0.3, (g, 2))
        gt = np.concatenate([c - wh / 2, c + wh / 2], 1) * 100.0
        gts.append(gt)
        k = 0 if i % 5 == 2 else int(r.randint(1, 40))
        sc = np.sort(r.randint(1, 12, k) / 12.0)[::-1].astype(np.float32)      # heavily tied scores
        for j in range(k):
            b = gt[r.randint(g)] / 100.0 if (g and r.rand() < 0.7) else r.uniform(0, 1, 4)   # exact copies: detected twice
            out[i, 1, j] = np.concatenate([[sc[j]], b])
    rows, offs = EA.collect_detections(out, 100.0, 100.0, 0.05)
    d_rows, d_off = collect_detections(torch.from_numpy(out).cuda(), 100.0, 100.0, 0.05, as_numpy=False)
    n = int(d_off[-1])
    assert np.array_equal(d_rows[:n].cpu().numpy(), rows)
    for metric in (True, False):
        o_ap, o_iobb, o_code, o_order = EA.ap_iobb(rows, gts, [0.5, 1.0], [0.5, 1.0], use_07_metric=metric)
        ap, iobb, tp, order = ap_iobb(d_rows, d_off, gts, [0.5, 1.0], [0.5, 1.0], use_07_metric=metric, details=True)
        assert np.array_equal(tp.cpu().numpy(), o_code) and np.array_equal(order.cpu().numpy().astype(np.int64), o_order)
        np.testing.assert_allclose(ap + iobb, o_ap + o_iobb, rtol=1e-12, atol=1e-15)
        assert (o_code == 1).any() and (o_code == 2).any() and (o_code == 0).any()


def test_validation_set_size_against_the_oracle():
    """1500 images x up to 200 detections (more than one radix tile per pass, several scan chunks)"""
    from grouped_ssd_pytorch_b200.layers.functions import ap_iobb, collect_detections
    r = np.random.RandomState(8)
    I, K = 1500, 200
    out = np.zeros((I, 2, K, 5), np.float32)
    gts = []
    for i in range(I):
        g = int(r.randint(1, 5))
        c = r.uniform(0.2, 0.8, (g, 2)); wh = r.uniform(0.05, 0.3, (g, 2))
        gt = np.concatenate([c - wh / 2, c + wh / 2], 1)
        gts.append(gt * 300.0)
        k = int(r.randint(0, K + 1))
        sc = np.sort(r.uniform(0.01, 1, k).astype(np.float32))[::-1]
        pick = r.randint(0, g, k)
        boxes = np.where(r.rand(k, 1) < 0.4, gt[pick] + r.randn(k, 4) * 0.03, r.uniform(0, 1, (k, 4)))
        out[i, 1, :k, 0] = sc
        out[i, 1, :k, 1:] = boxes
    rows, offs = EA.collect_detections(out, 300.0, 300.0, 0.2)
    d_rows, d_off = collect_detections(torch.from_numpy(out).cuda(), 300.0, 300.0, 0.2, as_numpy=False)
    assert int(d_off[-1]) == rows.shape[0] > 100000
    o_ap, o_iobb, o_code, o_order = EA.ap_iobb(rows, gts, [0.3, 0.5, 0.7], [0.3, 0.5, 0.7])
    ap, iobb, tp, order = ap_iobb(d_rows, d_off, gts, [0.3, 0.5, 0.7], [0.3, 0.5, 0.7], details=True)
    assert np.array_equal(tp.cpu().numpy(), o_code) and np.array_equal(order.cpu().numpy().astype(np.int64), o_order)
    np.testing.assert_allclose(ap + iobb, o_ap + o_iobb, rtol=1e-12, atol=1e-15)
