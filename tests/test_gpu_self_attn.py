"""GSSD++'s Self_Attn block with the attention core on the library's kernels (gssd_attn_fwd / gssd_attn_bwd):

* the drop-in module against the golden of the reference's OWN Self_Attn module (tests/golden/self_attn.npz, float64, eval
  mode): every output and every parameter gradient within 3e-3 of its scale (TF32 tensor-core products with fp32
  accumulation inside the kernels, fp32 torch convolutions around them);
* the core alone against the numpy oracle (oracle/self_attn.py) on shapes that exercise every kernel variant: tails in
  queries and keys, key counts that are not a multiple of 4, one key, one query, every channel-tile width, strips of 16 and 8
  queries (key lists beyond 2400), and the 38 x 38 map of GSSD++ (1444 x 1444) against torch's fp32 operators."""
import numpy as np
import pytest
import torch

import cases
from grouped_ssd_pytorch_b200 import _lib
from grouped_ssd_pytorch_b200.layers import self_attn as ours

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(DEV)


def rel(a, ref):
    ref = np.asarray(ref, np.float64)
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.mark.parametrize("tag", sorted(cases.SA_CASES))
def test_self_attn_module_vs_reference_golden(tag):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = cases.golden("self_attn")
    seed, B, C, H, factor = cases.SA_CASES[tag]
    x, prm, (u1, u2) = cases.sa_case(tag)
    m = ours.Self_Attn(C, factor).to(DEV).eval()
    m.load_state_dict({k: T(v) for k, v in prm.items()}, strict=True)            # the reference's parameter / buffer names
    xt = T(x).requires_grad_(True)
    n0 = _lib.launch_count()
    y, gated, attn = m(xt, True)
    ((y * T(u1)).sum() + (gated * T(u2)).sum()).backward()
    assert _lib.launch_count() >= n0 + 3, "forward + two backward kernels of the library"
    assert len(m(xt)) == 2                                                       # self_attn.py:88-89
    errs = dict(out=rel(y, g[tag + "_out"]), gated=rel(gated, g[tag + "_gated"]), attn=rel(attn, g[tag + "_attn"]),
                d_x=rel(xt.grad, g[tag + "_d_x"]), d_sigma=rel(m.sigma.grad, g[tag + "_d_sigma"]))
    for name in ("snconv1x1_theta", "snconv1x1_phi", "snconv1x1_g", "snconv1x1_attn"):
        mod = getattr(m, name)
        errs[name + ".w"] = rel(cases.strided_sample(mod.weight_orig.grad.cpu().numpy())[:-2], g["%s_d_%s_w_s" % (tag, name)][:-2])
        errs[name + ".b"] = rel(mod.bias.grad, g["%s_d_%s_b" % (tag, name)])
    # the bias of phi has NO gradient (shifting every key's score by the same amount leaves the softmax unchanged): the sum of the
    # gradient of phi over the keys cancels to rounding noise on both sides; measure it against the terms that cancel
    cancel = float(np.abs(g[tag + "_d_phi_conv"]).sum((0, 2, 3)).max())
    errs["snconv1x1_phi.b"] = float(np.abs(m.snconv1x1_phi.bias.grad.cpu().numpy() - g[tag + "_d_snconv1x1_phi_b"]).max() / max(cancel, 1e-30))
    print(tag, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 3e-3, errs


CORE_SHAPES = [
    # B, D, Cv, N, M
    (2, 32, 128, 25, 25), (1, 64, 256, 100, 100), (3, 32, 32, 37, 5), (2, 64, 64, 9, 1), (2, 32, 128, 1, 1), (1, 128, 512, 361, 361),
    (1, 32, 96, 70, 33), (1, 32, 128, 40, 2000), (1, 32, 128, 3000, 40), (1, 32, 128, 20, 4500), (1, 64, 256, 1444, 1444),
]


@pytest.mark.parametrize("shape", CORE_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_attention_core_vs_oracle(shape):
    from oracle import self_attn as SA
    B, D, Cv, N, M = shape
    r = np.random.RandomState(B * 7 + N + M)
    theta = (r.standard_normal((B, D, N)) * 0.6).astype(np.float32)
    phi = (r.standard_normal((B, D, M)) * 0.6).astype(np.float32)
    gg = r.standard_normal((B, Cv, M)).astype(np.float32)
    d_o = r.standard_normal((B, Cv, N)).astype(np.float32)
    tt, tp, tg = T(theta).requires_grad_(True), T(phi).requires_grad_(True), T(gg).requires_grad_(True)
    out, attn = ours.attention_core(tt, tp, tg)
    assert not attn.requires_grad
    out.backward(T(d_o))
    a_ref, o_ref = SA.attention(theta, phi, gg)
    d_theta, d_phi, d_g = SA.attention_backward(theta, phi, gg, a_ref, d_o)
    errs = dict(attn=rel(attn, a_ref), attn_g=rel(out, o_ref), d_theta=rel(tt.grad, d_theta), d_phi=rel(tp.grad, d_phi), d_g=rel(tg.grad, d_g))
    print(shape, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 3e-3, errs                  # TF32 operands (10-bit mantissa), fp32 accumulation and softmax
    rows = attn.sum(-1)
    assert float((rows - 1).abs().max()) <= 1e-5


def test_attention_core_argument_errors():
    t = torch.randn(1, 32, 8, device=DEV)
    with pytest.raises(NotImplementedError):
        ours.attention_core(torch.randn(1, 8, 8, device=DEV), torch.randn(1, 8, 8, device=DEV), torch.randn(1, 32, 8, device=DEV))   # D = 8
    with pytest.raises(ValueError):
        ours.attention_core(t, torch.randn(1, 32, 9, device=DEV), torch.randn(1, 32, 8, device=DEV))
    with pytest.raises(RuntimeError):
        ours.attention_core(t.cpu(), t.cpu(), t.cpu())
    lib = _lib.load()
    assert lib.gssd_attn_fwd(None, None, None, 1, 32, 32, 8, 8, None, None, None) == _lib.ERR_ARG
    assert lib.gssd_attn_fwd(t.data_ptr(), t.data_ptr(), t.data_ptr(), 1, 32, 32, 8, 100000, t.data_ptr(), t.data_ptr(), None) == _lib.ERR_LIMIT
