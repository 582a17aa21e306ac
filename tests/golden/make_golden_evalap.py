"""Generate tests/golden/evalap.npz: the reference's OWN evaluator (test_ap_iobb.py: test_net -> make_pred -> voc_ap) driven by a stub
network and a stub dataset that replay seeded Detect outputs and ground-truth boxes.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_evalap.py       (build container only: needs /root/reference)

Stores, per case: the Detect outputs [I,2,200,5], the image sizes, the ground truth, and what the reference returns
(ap_result, iobb_result) for both AP metrics."""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
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
import test_ap_iobb as E  # noqa: E402  (reference)

AP_LIST, IOBB_LIST = [0.3, 0.5, 0.7], [0.3, 0.5, 0.7]      # train_lesion_multiphase_v2.py: ap_list / iobb_list


def make_case(seed, n_img, n_det_max, W, H, jitter):
    """seeded Detect-shaped outputs: per image 0..n_det_max rows in descending score, some of them jittered copies of the
    ground-truth boxes (true positives), zero padded to 200; ground truth 0-4 boxes per image in pixel coordinates"""
    r = np.random.RandomState(seed)
    out = np.zeros((n_img, 2, 200, 5), np.float32)
    gts = []
    for i in range(n_img):
        g = int(r.randint(0 if seed % 2 else 1, 5))
        c = r.uniform(0.15, 0.85, (g, 2)); wh = r.uniform(0.05, 0.3, (g, 2))
        gt = np.concatenate([np.clip(c - wh / 2, 0, 1), np.clip(c + wh / 2, 0, 1)], 1)
        gts.append((gt * np.array([W, H, W, H])).astype(np.float64))          # annotation boxes are in pixels
        k = int(r.randint(0, n_det_max + 1))
        boxes = []
        for j in range(k):
            if g and r.rand() < 0.5:
                b = gt[r.randint(g)] + r.randn(4) * jitter
            else:
                cc = r.uniform(0.1, 0.9, 2); ww = r.uniform(0.05, 0.3, 2)
                b = np.concatenate([cc - ww / 2, cc + ww / 2])
            boxes.append(b)
        sc = np.sort(r.uniform(0.02, 1.0, k).astype(np.float32))[::-1]
        # (no equal scores here: the reference orders them with an unstable np.argsort, i.e. its result is not defined for ties;
        # the tie contract — equal scores keep (image, rank) order — is tested between the oracle and the kernels)
        assert np.unique(sc).size == k
        for j in range(k):
            out[i, 1, j, 0] = sc[j]
            out[i, 1, j, 1:] = boxes[j]
    return out, gts


class StubSet(object):
    def __init__(self, name, out, gts, W, H):
        self.name, self.out, self.gts, self.W, self.H = name, out, gts, W, H

    def __len__(self):
        return self.out.shape[0]

    def pull_image(self, idx):
        img = np.zeros((4, self.H, self.W, 3), np.float32)
        img[0, 0, 0, 0] = idx                                                  # the stub net reads the index back from the pixels
        return img

    def pull_anno(self, idx):
        g = self.gts[idx]
        return np.concatenate([g, np.zeros((g.shape[0], 1))], 1)


def main():
    out_npz = {}
    for tag, (seed, n_img, n_det_max, W, H, jitter, thresh) in {
            "a": (1, 40, 12, 512, 384, 0.02, 0.2), "b": (2, 64, 200, 300, 300, 0.04, 0.05), "c": (3, 7, 3, 640, 480, 0.01, 0.5)}.items():
        out, gts = make_case(seed, n_img, n_det_max, W, H, jitter)
        ds = StubSet("lesion_test_ap", out, gts, W, H)
        transform = lambda img: (img,)
        net = lambda x: torch.from_numpy(out[int(round(float(x[0, 0, 0, 0])))][None])
        for metric in (True, False):
            ap, iobb = E.test_net(net, False, ds, transform, 300, thresh=thresh, mode='v2', use_07_metric=metric,
                                  ap_list=AP_LIST, iobb_list=IOBB_LIST)
            out_npz["%s/ap_%d" % (tag, int(metric))] = np.asarray(ap, np.float64)
            out_npz["%s/iobb_%d" % (tag, int(metric))] = np.asarray(iobb, np.float64)
        out_npz[tag + "/out"] = out
        out_npz[tag + "/gt"] = np.concatenate(gts, 0) if gts else np.zeros((0, 4))
        out_npz[tag + "/gt_off"] = np.cumsum([0] + [g.shape[0] for g in gts]).astype(np.int32)
        out_npz[tag + "/meta"] = np.array([W, H, thresh], np.float64)
        print(tag, out_npz[tag + "/ap_1"], out_npz[tag + "/iobb_1"], out_npz[tag + "/ap_0"])
    path = os.path.join(HERE, "evalap.npz")
    np.savez_compressed(path, **out_npz)
    print("evalap %8.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
