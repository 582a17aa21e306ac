"""The selection algorithm of csrc/select.cuh (per-image radix select over the CTAs of a cluster, 64-bit cut, tie rules) as a
Python model, against its definition: "the k largest keys, equal keys taken by lower index first (loss: OHNM, multibox_loss.py:
102-106 with the oracle's stable rule) or by higher index first (Detect / nms top-k, box_utils.py:194-196)".

This checks the ALGORITHM the kernels implement — digit search on the summed histograms, the one-warp ranking of a bin with at
most 32 members, and the path for more than 32 exactly equal keys at the cut with its per-slice `before / want` bookkeeping — on
inputs with heavy ties that the GPU parity tests only reach for a single CTA per image.  The CUDA code itself is covered by
tests/test_gpu_parity.py."""
import numpy as np
import pytest

MASK32 = 0xFFFFFFFF


def composite(key, index, low_first):
    return (int(key) << 32) | ((~int(index)) & MASK32 if low_first else int(index))


def model_select(slices, k, low_first):
    """slices: list of uint32 arrays (contiguous slices of one image, in rank order) -> list of boolean masks"""
    bases = np.cumsum([0] + [len(s) for s in slices])[:-1]
    prefix = mask = 0
    k_rem, eq_total, digit = k, 0, 0
    hists = None
    for shift in (24, 16, 8, 0):
        hists = []
        for s in slices:
            sel = s[(s & mask) == prefix] if mask else s
            hists.append(np.bincount((sel >> shift) & 255, minlength=256))
        tot = np.sum(hists, axis=0)
        above = 0
        for d in range(255, -1, -1):                              # suffix sums, bin 255 first
            if above < k_rem <= above + tot[d]:
                digit, k_rem, eq_total = d, k_rem - above, int(tot[d])
                break
            above += tot[d]
        prefix |= digit << shift
        mask |= 255 << shift
        if eq_total <= 32:
            break
    if eq_total <= 32:
        members = [composite(key, b + i, low_first) for s, b in zip(slices, bases) for i, key in enumerate(s) if (int(key) & mask) == prefix]
        assert len(members) == eq_total
        cut = sorted(members, reverse=True)[k_rem - 1]
        cuts = [cut] * len(slices)
    else:
        v, need = prefix, k_rem
        take_all, take_none = v << 32, (v << 32) + (1 << 32)
        cuts = []
        for r, (s, b) in enumerate(zip(slices, bases)):
            if need == eq_total:
                cuts.append(take_all)
                continue
            eq_local = int(hists[r][digit])
            before = int(sum(h[digit] for h in hists[:r]))
            want = need - before if low_first else need - (eq_total - before - eq_local)
            if want <= 0:
                cuts.append(take_none)
            elif want >= eq_local:
                cuts.append(take_all)
            else:
                idx = np.flatnonzero(s == v)
                tie = idx[want - 1] if low_first else idx[len(idx) - want]
                cuts.append(composite(v, b + tie, low_first))
    return [np.array([composite(key, b + i, low_first) >= c for i, key in enumerate(s)], bool) for s, b, c in zip(slices, bases, cuts)]


def definition(keys, k, low_first):
    order = sorted(range(len(keys)), key=lambda i: (-int(keys[i]), i if low_first else -i))
    out = np.zeros(len(keys), bool)
    out[order[:k]] = True
    return out


def cases():
    r = np.random.RandomState(3)
    for trial in range(60):
        n = int(r.choice([5, 40, 300, 1500]))
        kind = trial % 5
        if kind == 0:
            keys = r.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)                  # no ties
        elif kind == 1:
            keys = r.choice(np.array([7, 0x80000000, 0x80000001, 0xBF800000], np.uint32), n)    # four values: huge ties
        elif kind == 2:
            keys = np.full(n, 0x3F800000, np.uint32)                                            # every key identical
        elif kind == 3:
            keys = (0x3F800000 + r.randint(0, 3, n)).astype(np.uint32)                          # ties in the last digit
        else:
            keys = (r.randint(0, 6, n).astype(np.uint32) << 16) | 0x40000000                    # ties resolved in pass 2
        nranks = int(r.choice([1, 2, 4, 8]))
        bounds = np.sort(r.randint(0, n + 1, nranks - 1))
        slices = np.split(keys, bounds)
        for k in sorted({1, max(1, n // 3), max(1, n - 1), int(r.randint(1, n + 1))}):
            yield keys, slices, k


@pytest.mark.parametrize("low_first", [True, False])
def test_select_model_matches_its_definition(low_first):
    n_checked = 0
    for keys, slices, k in cases():
        got = np.concatenate(model_select(slices, k, low_first)) if len(keys) else np.zeros(0, bool)
        assert got.sum() == k
        assert np.array_equal(got, definition(keys, k, low_first)), (len(keys), k, [len(s) for s in slices])
        n_checked += 1
    assert n_checked > 150
