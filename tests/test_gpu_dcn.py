"""GSSD++'s modulated deformable convolution on the library's kernels (gssd_dcn_columns + gssd_conv_igemm forward;
gssd_conv_igemm / gssd_conv_wgrad / gssd_dcn_columns_bwd backward) through the drop-in `layers.dcn_v2_custom.DCN`:

* against the golden of the reference's OWN DCN module (tests/golden/dcn.npz, float64, operator = torchvision's
  deform_conv2d) within the north-star tolerance of the bf16 convolution path, 1e-2 of each tensor's scale;
* against the numpy oracle (oracle/dcn.py) fed the bf16-rounded operands the kernels see: within bf16 rounding of the result;
* at the size GSSD++ runs it (1024 -> 512 channels, 38 x 38, 4 deformable groups; ssd_multiphase_custom_group.py:164-169)
  against torchvision's fp32 CUDA operator."""
import numpy as np
import pytest
import torch

import cases
from grouped_ssd_pytorch_b200 import _lib
from grouped_ssd_pytorch_b200.layers import dcn_v2_custom as ours

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(DEV)


def rel(a, ref):
    ref = np.asarray(ref, np.float64)
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    assert a.shape == ref.shape
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def run_module(tag):
    seed, N, C, O, H, W, dg, s = cases.DCN_CASES[tag]
    c = cases.dcn_case(tag)
    m = ours.DCN(C, O, kernel_size=3, stride=1, padding=1, deformable_groups=dg).to(DEV)
    with torch.no_grad():
        m.weight.copy_(T(c["weight"])); m.bias.copy_(T(c["bias"]))
        m.conv_offset_mask.weight.copy_(T(c["com_w"])); m.conv_offset_mask.bias.copy_(T(c["com_b"]))
    x = T(c["x"]).requires_grad_(True)
    kept = {}

    def keep(mod, inputs, o):
        o.retain_grad()
        kept["om"] = o

    h = m.conv_offset_mask.register_forward_hook(keep)
    y, offset = m(x)
    h.remove()
    (y * T(c["gout"])).sum().backward()
    return m, x, y, offset, kept["om"]


@pytest.mark.parametrize("tag", sorted(cases.DCN_CASES))
def test_dcn_module_vs_reference_golden(tag):
    torch.backends.cudnn.allow_tf32 = False
    g = cases.golden("dcn")
    n0 = _lib.launch_count()
    m, x, y, offset, om = run_module(tag)
    assert _lib.launch_count() > n0, "the CUDA library did not launch"
    assert rel(offset, g[tag + "_offset"]) <= 1e-5                     # the offset convolution is torch's own fp32 conv
    errs = dict(out=rel(y, g[tag + "_out"]), d_input=rel(x.grad, g[tag + "_d_input"]), d_om=rel(om.grad, g[tag + "_d_om"]),
                d_bias=rel(m.bias.grad, g[tag + "_d_bias"]),
                d_weight=rel(cases.strided_sample(m.weight.grad.cpu().numpy())[:-2], g[tag + "_d_weight_s"][:-2]))
    print(tag, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 1e-2, errs


def test_dcn_operator_vs_oracle_on_the_operands_the_kernels_see():
    """bf16-rounded input and filter, columns rounded to bf16, fp32 accumulation: what is left is the rounding of the
    bf16 output (2^-9 relative) and of d_columns."""
    from oracle import dcn as D
    from oracle.source_block import bf16_round
    c = cases.dcn_case("rand")
    g = cases.golden("dcn")
    om = g["rand_om"].astype(np.float64)
    k = om.shape[1] // 3
    offset, mask = om[:, :2 * k].astype(np.float32), (1 / (1 + np.exp(-om[:, 2 * k:]))).astype(np.float32)
    xb, wb = bf16_round(c["x"]), bf16_round(c["weight"])
    col = bf16_round(D.columns(xb, offset, mask, c["dg"]).astype(np.float32))
    w2 = wb.astype(np.float64).reshape(wb.shape[0], wb.shape[1], 9).transpose(0, 2, 1)
    out_ref = np.einsum("ntchw,otc->nohw", col.astype(np.float64), w2) + c["bias"].astype(np.float64)[None, :, None, None]
    xt, ot, mt = T(c["x"]).requires_grad_(True), T(offset).requires_grad_(True), T(mask).requires_grad_(True)
    wt, bt = T(c["weight"]).requires_grad_(True), T(c["bias"]).requires_grad_(True)
    y = ours.dcn_v2_conv(xt, ot, mt, wt, bt, 1, 1, 1, c["dg"])
    assert rel(y, out_ref) <= 4e-3
    gy = bf16_round(c["gout"])
    (y * T(gy)).sum().backward()
    b = D.backward(xb, offset, mask, wb, c["dg"], gy)
    errs = dict(d_input=rel(xt.grad, b["d_input"]), d_offset=rel(ot.grad, b["d_offset"]), d_mask=rel(mt.grad, b["d_mask"]),
                d_weight=rel(wt.grad, b["d_weight"]), d_bias=rel(bt.grad, b["d_bias"]))
    print({k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 6e-3, errs


def test_dcn_at_the_gssdpp_size_vs_torchvision_fp32():
    from torchvision.ops import deform_conv2d
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device=DEV).manual_seed(7)
    N, C, O, H, W, dg = 2, 1024, 512, 38, 38, 4
    rn = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    x, w, b = rn(N, C, H, W), rn(O, C, 3, 3) / (9 * C) ** 0.5, 0.1 * rn(O)
    off, msk = 1.5 * rn(N, 2 * dg * 9, H, W), torch.sigmoid(rn(N, dg * 9, H, W))
    gout = rn(N, O, H, W)
    res = []
    for fn in (lambda *a: ours.dcn_v2_conv(*a, 1, 1, 1, dg),
               lambda xi, oi, mi, wi, bi: deform_conv2d(xi, oi, wi, bi, stride=1, padding=1, dilation=1, mask=mi)):
        leaves = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
        y = fn(*leaves)
        (y * gout).sum().backward()
        res.append([y] + [t.grad for t in leaves])
    names = ("out", "d_input", "d_offset", "d_mask", "d_weight", "d_bias")
    errs = {n: rel(a, r.detach().double().cpu().numpy()) for n, a, r in zip(names, res[0], res[1])}
    print({k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 1e-2, errs


def test_dcn_argument_errors():
    x = torch.randn(1, 128, 5, 5, device=DEV)
    w = torch.randn(64, 128, 3, 3, device=DEV)
    off, msk = torch.zeros(1, 18, 5, 5, device=DEV), torch.ones(1, 9, 5, 5, device=DEV)
    with pytest.raises(NotImplementedError):
        ours.dcn_v2_conv(x, off, msk, w, None, 2, 1, 1, 1)                   # stride 2
    with pytest.raises(NotImplementedError):
        ours.dcn_v2_conv(x[:, :96], off, msk, w[:, :96].contiguous(), None, 1, 1, 1, 1)     # c_in not a multiple of 128
    with pytest.raises(ValueError):
        ours.dcn_v2_conv(x, off[:, :16], msk, w, None, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):
        ours.dcn_v2_conv(x.cpu(), off.cpu(), msk.cpu(), w.cpu(), None, 1, 1, 1, 1)
    lib = _lib.load()
    assert lib.gssd_dcn_columns(None, None, None, 1, 128, 5, 5, 1, None, None) == _lib.ERR_ARG
    assert lib.gssd_dcn_columns(x.data_ptr(), off.data_ptr(), msk.data_ptr(), 1, 100, 5, 5, 3, x.data_ptr(), None) == _lib.ERR_ARG
    assert lib.gssd_dcn_columns(x.data_ptr(), off.data_ptr(), msk.data_ptr(), 1, 100, 5, 5, 5, x.data_ptr(), None) == _lib.ERR_LIMIT
