"""GPU: the C-ABI packers of the tensor-core operand images (lrds_pack_gmm_mix_tc, lrds_pack_logreg_tc) agree with the
host-side packers the Python classes use (whose byte layout tests/test_capi_symbols.py pins on the CPU)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _decode(img, part, unscale_at=None):
    hi = img[..., :part].contiguous().view(torch.float16).float()
    lo = img[..., part:2 * part].contiguous().view(torch.float16).float()
    at = 2 * part if unscale_at is None else unscale_at
    unscale = img[..., at:at + 4].contiguous().view(torch.float32)
    return (hi + lo) * unscale


@pytest.mark.parametrize("shared", [False, True])
@pytest.mark.parametrize("steps", [1, 5])
def test_gmm_mixture_image_packer_matches_host_packer(device, steps, shared):
    from sde_sampler_lrds_b200 import _native as N
    from sde_sampler_lrds_b200.distr.base import fill_gmm, gmm_block
    torch.manual_seed(3)
    M, d = 7, 13
    lead = (steps,) if steps > 1 else ()
    loc, var = torch.randn(*lead, M, d) * 6, torch.rand(*lead, M, d) * 3 + 0.02
    if shared:
        var = var[..., :1, :].expand_as(loc).contiguous()
    block = gmm_block(loc, var, torch.rand(M) + 0.1, device)
    g = fill_gmm(N.Gmm(), block, stepped=steps > 1)
    d_pad = 16
    nbytes = N.lib().lrds_gmm_mix_tc_bytes(M, d_pad)
    want = block[4].reshape(steps, -1)
    assert want.shape[1] == nbytes
    out = torch.zeros(steps, nbytes, dtype=torch.uint8, device=device)
    N.check(N.lib().lrds_pack_gmm_mix_tc(C.byref(g), d, d_pad, steps, C.c_void_p(out.data_ptr()), N.stream_ptr(device)))
    torch.cuda.synchronize()
    Mp, Kin = 16, 16 * ((d_pad + 15) // 16)
    lpart = Mp * Kin * 2
    contr = nbytes - (2 * lpart + 4 * Mp + 16)
    part = (contr - 16) // 2
    a, b = _decode(out.cpu(), part), _decode(want.cpu(), part)
    assert torch.equal(out[:, 2 * part:2 * part + 4].cpu(), want[:, 2 * part:2 * part + 4].cpu())  # same power of two
    assert ((a - b).abs() <= 2.0 ** -20 * b.abs().clamp(min=2.0 ** -10)).all()
    # the logit image: same power of two, same matrix, same c_m / norms / shared flag
    lo_, lw = out[:, contr:].cpu(), want[:, contr:].cpu()
    tail_o, tail_w = lo_[:, 2 * lpart + 4 * Mp:].contiguous().view(torch.float32), lw[:, 2 * lpart + 4 * Mp:].contiguous().view(torch.float32)
    assert torch.equal(tail_o[:, 0], tail_w[:, 0]) and torch.equal(tail_o[:, 3], tail_w[:, 3])
    assert (tail_o[:, 3] == (1.0 if shared else 0.0)).all()
    assert ((tail_o[:, 1:3] - tail_w[:, 1:3]).abs() <= 1e-5 * tail_w[:, 1:3].abs()).all()
    a, b = _decode(lo_, lpart, 2 * lpart + 4 * Mp), _decode(lw, lpart, 2 * lpart + 4 * Mp)
    assert ((a - b).abs() <= 2.0 ** -18 * b.abs().max()).all()
    co, cw = lo_[:, 2 * lpart:2 * lpart + 4 * Mp].contiguous().view(torch.float32), lw[:, 2 * lpart:2 * lpart + 4 * Mp].contiguous().view(torch.float32)
    assert torch.equal(torch.isinf(co), torch.isinf(cw))
    fin = torch.isfinite(cw)
    assert ((co[fin] - cw[fin]).abs() <= 2e-6 * cw[fin].abs().clamp(min=1.0)).all()


def test_logreg_image_packer_matches_host_packer(device):
    from sde_sampler_lrds_b200 import _native as N
    from sde_sampler_lrds_b200.distr.logistic_regression import LogisticRegression
    torch.manual_seed(4)
    n, p = 37, 21
    X, y = torch.rand(n, p) * 0.7, (torch.rand(n) > 0.5).float()
    tgt = LogisticRegression(X_train=X, y_train=y)
    distr, keep = tgt._lrds_pack(device)
    want = keep[3]
    nbytes = N.lib().lrds_logreg_tc_bytes(n, p)
    assert want.numel() == nbytes
    out = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    d_pad = ((p + 1 + 7) // 8) * 8
    N.check(N.lib().lrds_pack_logreg_tc(C.byref(distr.logreg), d_pad, C.c_void_p(out.data_ptr()), N.stream_ptr(device)))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), want.cpu())
