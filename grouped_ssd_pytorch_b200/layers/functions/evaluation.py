"""The consumer of Detect's output in the reference's evaluator (test_ap_iobb.py), on the GPU and for whole batches:

  collect_detections(output, width, height, thresh)  make_pred's per-image post-filter (test_ap_iobb.py:122-149)
  ap_iobb(rows, gt_list, ap_list, iobb_list)          test_net's AP / IoBB computation (test_ap_iobb.py:243-326 + voc_ap 10-41)
  evaluate_detections(outputs, sizes, gt_list, ...)   both, for the Detect outputs of a whole validation set

The reference runs this one image at a time on the host ("sluggish cpu-bound", train_lesion_multiphase_v2.py:111); here the
filter is a prefix count + compaction, the TP / FP assignment one warp per image, the global score order a radix sort and the
precision / recall curve a scan, all in libgssd_b200.so (csrc/evalap.cu).  No CPU fallback."""
import ctypes as C

import numpy as np
import torch

from ... import _lib


def collect_detections(output, width, height, thresh, class_index=1, first_image_id=0, as_numpy=True):
    """output: Detect's [B, C, top_k, 5] -> rows [n, 6] = (image id, score, xmin, ymin, xmax, ymax) in image order, descending
    score inside an image — what the reference accumulates in `predictions` (test_ap_iobb.py:124-153).  as_numpy=False returns
    the device tensors (rows, offsets[B+1]) without any host synchronisation."""
    lib = _lib.require_cuda()
    dev = _lib.device_of(output)
    with torch.cuda.device(dev):
        out = _lib.f32(output, dev)
        B, Cn, K = out.shape[0], out.shape[1], out.shape[2]
        rows = torch.empty((B * K, 6), dtype=torch.float32, device=dev)
        offsets = torch.empty((B + 1,), dtype=torch.int32, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        _lib.check(lib.gssd_collect_detections(out.data_ptr(), B, Cn, K, int(class_index), float(width), float(height), float(thresh),
                                               int(first_image_id), rows.data_ptr(), offsets.data_ptr(), counts.data_ptr(),
                                               _lib.stream()), "gssd_collect_detections")
    if not as_numpy:
        return rows, offsets
    n = int(offsets[-1])                                             # the one synchronisation: the caller wants a host array
    return rows[:n].cpu().numpy()


def ap_iobb(rows, det_offsets, gt_list, ap_list, iobb_list, use_07_metric=True, details=False):
    """AP at the IoU thresholds `ap_list` and at the IoBB thresholds `iobb_list` -> (ap_result, iobb_result), the lists
    test_net returns (test_ap_iobb.py:231-328).
    rows: [n, 6] device tensor grouped by image (collect_detections), det_offsets: int32 [I+1] first row of every image;
    gt_list: per image an [G_i, 4] (or [G_i, 5]: the label column is dropped, test_ap_iobb.py:194) array of boxes in the rows'
    coordinates.  Equal scores keep (image, rank) order (the reference's unstable argsort leaves it undefined)."""
    lib = _lib.require_cuda()
    dev = rows.device
    n_img = len(gt_list)
    n_iou, n_iobb = len(ap_list), len(iobb_list)
    gts = [np.asarray(g, np.float32).reshape(-1, np.asarray(g).shape[-1] if np.asarray(g).ndim == 2 else 4)[:, :4] for g in gt_list]
    npos = int(sum(g.shape[0] for g in gts))
    n_det = int(det_offsets[-1])
    if n_det == 0 or npos == 0:                                      # test_ap_iobb.py:243-249: nothing predicted
        return ([0.] * n_iou, [0.] * n_iobb) + ((None, None) if details else ())
    if max(g.shape[0] for g in gts) > 128:
        raise RuntimeError("ap_iobb: more than 128 ground-truth boxes in one image")
    gt_off = np.zeros(n_img + 1, np.int32)
    np.cumsum([g.shape[0] for g in gts], out=gt_off[1:])
    with torch.cuda.device(dev):
        gt = torch.from_numpy(np.concatenate(gts, 0) if npos else np.zeros((0, 4), np.float32)).to(dev)
        gto = torch.from_numpy(gt_off).to(dev)
        thr = torch.tensor([float(t) for t in ap_list] + [float(t) for t in iobb_list], dtype=torch.float64, device=dev)
        pts = torch.from_numpy(np.arange(0., 1.1, 0.1)).to(dev)                      # voc_ap: for t in np.arange(0., 1.1, 0.1)
        ap = torch.empty((n_iou + n_iobb,), dtype=torch.float64, device=dev)
        wsb = lib.gssd_ap_workspace_bytes(n_det, n_iou + n_iobb)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        tp = torch.empty((n_iou + n_iobb, n_det), dtype=torch.uint8, device=dev) if details else None
        order = torch.empty((n_det,), dtype=torch.int32, device=dev) if details else None
        _lib.check(lib.gssd_ap_eval(rows.data_ptr(), det_offsets.data_ptr(), gt.data_ptr(), gto.data_ptr(), n_img, n_det,
                                    thr.data_ptr(), n_iou, n_iobb, npos, 1 if use_07_metric else 0, pts.data_ptr(), ap.data_ptr(),
                                    _lib.ptr(tp), _lib.ptr(order), ws.data_ptr(), wsb, _lib.stream()), "gssd_ap_eval")
        res = ap.cpu().numpy().tolist()
    out = (res[:n_iou], res[n_iou:])
    return out + ((tp, order) if details else ())


def evaluate_detections(outputs, sizes, gt_list, thresh=0.05, ap_list=(0.5,), iobb_list=(0.1,), use_07_metric=True, class_index=1):
    """test_net for Detect outputs that are already computed (batched inference): outputs = list of [B_k, C, top_k, 5] tensors
    covering the images 0..I-1 in order, sizes = (width, height) of the images or a list of one pair per output batch,
    gt_list = per image [G_i, 4|5] boxes in pixels -> (ap_result, iobb_result)."""
    rows, offs, first = [], [], 0
    for k, out in enumerate(outputs):
        w, h = sizes[k] if isinstance(sizes[0], (tuple, list)) else sizes
        r, o = collect_detections(out, w, h, thresh, class_index=class_index, first_image_id=first, as_numpy=False)
        rows.append(r); offs.append(o)
        first += out.shape[0]
    # stitch the batches: rows are compacted per batch, so gather the valid prefixes (sizes are on the device: one sync here)
    ns = [int(o[-1]) for o in offs]
    rows = torch.cat([r[:n] for r, n in zip(rows, ns)], 0)
    base, parts = 0, []
    for o, n in zip(offs, ns):
        parts.append(o[:-1] + base)
        base += n
    det_off = torch.cat(parts + [torch.tensor([base], dtype=torch.int32, device=rows.device)], 0).to(torch.int32)
    return ap_iobb(rows, det_off, gt_list, list(ap_list), list(iobb_list), use_07_metric=use_07_metric)
