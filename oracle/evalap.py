"""numpy restatement of the reference's evaluator (test_ap_iobb.py) — TEST INFRASTRUCTURE ONLY.

  collect_detections : make_pred, test_ap_iobb.py:122-149 (class-1 slab, score > 0, boxes * scale, image id, score > thresh)
  ap_iobb            : make_pred 213-223 (global descending-score order) + test_net 243-326 + voc_ap 10-41

Pinned by tests/golden/make_golden_evalap.py, which drives the reference's OWN test_net with a stub net / dataset that replay
seeded Detect outputs (tests/golden/evalap.npz).  One stated difference: the reference orders equal scores with an unstable
np.argsort; here (and in the kernels) equal scores keep (image, rank) order."""
import numpy as np


def collect_detections(output, width, height, thresh, class_index=1, first_image_id=0):
    """output [B, C, top_k, 5] -> rows [n, 6] float32 (image id, score, x1, y1, x2, y2), and offsets [B+1]"""
    out = np.asarray(output, np.float32)
    scale = np.array([width, height, width, height], np.float32)               # test_ap_iobb.py:127-128
    rows, offs = [], [0]
    for b in range(out.shape[0]):
        det = out[b, class_index]                                              # 131
        det = det[det[:, 0] > 0]                                               # 132-133
        boxes = np.hstack([np.full((det.shape[0], 1), first_image_id + b, np.float32), det[:, :1], det[:, 1:] * scale])   # 138-144
        boxes = boxes[boxes[:, 1] > np.float32(thresh)]                        # 148
        rows.append(boxes.astype(np.float32))
        offs.append(offs[-1] + boxes.shape[0])
    return (np.concatenate(rows, 0) if rows else np.zeros((0, 6), np.float32)), np.asarray(offs, np.int32)


def voc_ap(rec, prec, use_07_metric=True):
    """test_ap_iobb.py:10-41"""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def ap_iobb(rows, gt_list, ap_list, iobb_list, use_07_metric=True):
    """rows [n, 6] grouped by image id (ids index gt_list); gt_list: per image [G_i, 4] boxes in the rows' coordinates.
    -> (ap_result, iobb_result, tp codes [n_thr, n] in the rows' order (1 TP, 2 FP, 0 neither), order [n])"""
    rows = np.asarray(rows, np.float32)
    n = rows.shape[0]
    npos = sum(int(np.asarray(g).shape[0]) for g in gt_list)                   # mode 'v2', 195-198
    thr = [float(t) for t in ap_list] + [float(t) for t in iobb_list]
    n_iou = len(ap_list)
    order = np.argsort(-rows[:, 1].astype(np.float64), kind="stable")          # 213-223 (np.argsort there is unstable on ties)
    det = [[np.zeros(np.asarray(g).shape[0], bool) for g in gt_list] for _ in thr]
    code = np.zeros((len(thr), n), np.uint8)
    for d in order:                                                            # 251
        img = int(rows[d, 0])
        BBGT = np.asarray(gt_list[img], np.float64).reshape(-1, 4)
        bb = rows[d, 2:].astype(float)
        if BBGT.size > 0:
            ixmin = np.maximum(BBGT[:, 0], bb[0]); iymin = np.maximum(BBGT[:, 1], bb[1])
            ixmax = np.minimum(BBGT[:, 2], bb[2]); iymax = np.minimum(BBGT[:, 3], bb[3])
            iw = np.maximum(ixmax - ixmin, 0.); ih = np.maximum(iymax - iymin, 0.)
            inters = iw * ih
            uni_iou = ((bb[2] - bb[0]) * (bb[3] - bb[1]) + (BBGT[:, 2] - BBGT[:, 0]) * (BBGT[:, 3] - BBGT[:, 1]) - inters)
            uni_iobb = (bb[2] - bb[0]) * (bb[3] - bb[1])
            ov = (inters / uni_iou, inters / uni_iobb)
            for t, th in enumerate(thr):
                o = ov[0] if t < n_iou else ov[1]
                if np.max(o) > th:
                    j = int(np.argmax(o))
                    if not det[t][img][j]:
                        code[t, d] = 1; det[t][img][j] = True
                    else:
                        code[t, d] = 2
                else:
                    code[t, d] = 2
    res = []
    for t in range(len(thr)):
        c = code[t, order]
        tp, fp = np.cumsum(c == 1).astype(np.float64), np.cumsum(c == 2).astype(np.float64)
        rec = tp / float(npos)
        prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        res.append(float(voc_ap(rec, prec, use_07_metric)) if n else 0.0)
    return res[:n_iou], res[n_iou:], code, order
