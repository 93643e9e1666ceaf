"""GPU: each kernel of the hot path against the CPU oracle on the same seeded inputs (through the C ABI)."""
import ctypes

import pytest
import torch

from oracle import kd_oracle as O
from oracle import manual_backward as M
from tests.harness import relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from imagecaptioner_b200 import _ops
    _ops.load_library()
    return _ops


def _ref_gemm(A, B, a_mn, b_mn, M_, N_, K_):
    a = A.float().cpu()
    b = B.float().cpu()
    a = a[:K_, :M_].t() if a_mn else a[:M_, :K_]
    b = b[:K_, :N_].t() if b_mn else b[:N_, :K_]
    return a.double() @ b.double().t()


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
def test_gemm_fp32_ffma(a_mn, b_mn):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    for (M_, N_, K_) in [(70, 45, 33), (128, 64, 64), (257, 130, 100)]:
        A = torch.randn((K_, M_) if a_mn else (M_, K_), generator=g).to(DEV)
        B = torch.randn((K_, N_) if b_mn else (N_, K_), generator=g).to(DEV)
        bias = torch.randn(N_, generator=g).to(DEV)
        C = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, torch.float32, bias=bias, relu=True)
        ref = (_ref_gemm(A, B, a_mn, b_mn, M_, N_, K_) + bias.cpu().double()).clamp_min(0)
        assert relerr(C.cpu(), ref) < 1e-5, (M_, N_, K_)
        C0 = torch.randn(M_, N_, generator=g).to(DEV)
        C1 = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, alpha=0.5, beta=1.0, C=C0.clone())
        assert relerr(C1.cpu(), 0.5 * _ref_gemm(A, B, a_mn, b_mn, M_, N_, K_) + C0.cpu().double()) < 1e-5


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_bf16_tcgen05(a_mn, b_mn, out_dtype):
    """tcgen05/TMEM/TMA tile against fp64 matmul of the same bf16 operands: full tiles, ragged M/N/K, padded pitches."""
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    shapes = [(128, 128, 64), (512, 2048, 768), (200, 72, 40), (1000, 264, 520), (256, 512, 10240), (96, 5000, 256),
              (2048, 2048, 128), (4096, 2560, 192)]      # persistent: >1 tile per CTA (BN=128) and the BN=256 wide-output tile
    for (M_, N_, K_) in shapes:
        pad_a, pad_b = 8, 16                                  # pitches larger than the logical extent
        A = torch.randn((K_, M_ + pad_a) if a_mn else (M_, K_ + pad_a), generator=g).to(DEV).bfloat16()
        B = torch.randn((K_, N_ + pad_b) if b_mn else (N_, K_ + pad_b), generator=g).to(DEV).bfloat16()
        bias = torch.randn(N_, generator=g).to(DEV)
        C = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, out_dtype, bias=bias)
        ref = _ref_gemm(A, B, a_mn, b_mn, M_, N_, K_) + bias.cpu().double()
        tol = 3e-5 if out_dtype == torch.float32 else 6e-3    # fp32 accumulate over up to 10240 terms; bf16 output rounding 2^-8
        assert relerr(C.cpu(), ref) < tol, (M_, N_, K_, a_mn, b_mn)
        Cs = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, out_dtype, bias=bias, impl=1)    # FFMA tiles, same operands
        assert relerr(Cs.cpu(), ref) < tol
    # accumulate-into-C path (beta = 1) and ReLU
    M_, N_, K_ = 512, 256, 256
    A = torch.randn((K_, M_) if a_mn else (M_, K_), generator=g).to(DEV).bfloat16()
    B = torch.randn((K_, N_) if b_mn else (N_, K_), generator=g).to(DEV).bfloat16()
    C0 = torch.randn(M_, N_, generator=g).to(DEV).to(out_dtype)
    C1 = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, beta=1.0, C=C0.clone())
    tol = 3e-5 if out_dtype == torch.float32 else 1e-2     # bf16 beta path: accumulator and sum are each rounded to bf16
    assert relerr(C1.cpu(), _ref_gemm(A, B, a_mn, b_mn, M_, N_, K_) + C0.cpu().double()) < tol
    C2 = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, out_dtype, relu=True)
    assert relerr(C2.cpu(), _ref_gemm(A, B, a_mn, b_mn, M_, N_, K_).clamp_min(0)) < tol
    if out_dtype == torch.float32:
        # split-K path (skinny output, long reduction): accumulate onto C (beta = 1) and a strided C view (ldc > N, beta = 0)
        M_, N_, K_ = 256, 200, 4096
        A = torch.randn((K_, M_) if a_mn else (M_, K_), generator=g).to(DEV).bfloat16()
        B = torch.randn((K_, N_) if b_mn else (N_, K_), generator=g).to(DEV).bfloat16()
        ref = _ref_gemm(A, B, a_mn, b_mn, M_, N_, K_)
        C0 = torch.randn(M_, N_, generator=g).to(DEV)
        C1 = ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, beta=1.0, C=C0.clone())
        assert relerr(C1.cpu(), ref + C0.cpu().double()) < 3e-5
        wide = torch.full((M_, N_ + 56), 7.0, device=DEV)
        ops.gemm(A, B, M_, N_, K_, a_mn, b_mn, C=wide[:, 8:8 + N_])
        assert relerr(wide[:, 8:8 + N_].cpu(), ref) < 3e-5
        assert float(wide[:, :8].min()) == 7.0 and float(wide[:, 8 + N_:].min()) == 7.0      # neighbours untouched


def _token_loss(y, z, tgt, temp, alpha, w_ce, dtype):
    """kernel (3) through the C ABI -> (kd, ce, dlogits)"""
    ops = _ops()
    lib = ops.load_library()
    T, B, V = y.shape
    N = T * B
    yd = y.to(DEV).to(dtype).contiguous()
    zd = z.to(DEV).float().contiguous()
    td = tgt.to(DEV).contiguous()
    nval = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.b2c_count_valid(td.data_ptr(), N, V, nval.data_ptr(), st) == 0
    dy = torch.empty_like(yd)
    rows = torch.empty(2, N, device=DEV)
    rc = lib.b2c_kd_token_loss(yd.data_ptr(), zd.data_ptr(), td.data_ptr(), N, V, temp, alpha, w_ce, 1.0, nval.data_ptr(),
                               dy.data_ptr(), rows[0].data_ptr(), rows[1].data_ptr(), ops.dtype_code(dtype), st)
    assert rc == 0, lib.b2c_last_error()
    out5 = torch.empty(5, device=DEV)
    rc = lib.b2c_loss_finalize(rows[0].data_ptr(), rows[1].data_ptr(), N, nval.data_ptr(), 1.0, None, B, 1, None, 0, 1,
                               temp, alpha, 0.0, 0.0, w_ce, out5.data_ptr(), st)
    assert rc == 0, lib.b2c_last_error()
    torch.cuda.synchronize()
    return out5.cpu(), dy.float().cpu(), int(nval.item()), yd.float().cpu()


@pytest.mark.parametrize("V", [5000, 104, 203, 57])
@pytest.mark.parametrize("temp", [4.0, 2.0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kd_token_loss_kernel(V, temp, dtype):
    g = torch.Generator().manual_seed(V)
    T, B = 7, 9
    y = torch.randn(T, B, V, generator=g) * 1.5
    z = torch.randn(T, B, V, generator=g) * 2.0
    tgt = torch.randint(1, V, (T, B), generator=g)
    tgt[-2:, ::2] = 0                                        # PAD rows
    alpha, w_ce = 0.6, 0.25
    out5, dy, nval, y_used = _token_loss(y, z, tgt, temp, alpha, w_ce, dtype)
    assert nval == int((tgt != 0).sum())
    yy = y_used.double()                                     # the kernel's input after the dtype cast
    kd = O.token_kd(yy, z.double(), temp)
    ce = O.cross_entropy_ignore_pad(yy, tgt)
    assert abs(out5[2].item() - kd.item()) < 1e-4 * abs(kd.item()) + 1e-7
    assert abs(out5[1].item() - ce.item()) < 1e-4 * abs(ce.item())
    assert abs(out5[0].item() - (alpha * kd + w_ce * ce).item()) < 1e-4 * abs((alpha * kd + w_ce * ce).item())
    gref = M.kd_token_grad(yy, z.double(), tgt, temp, alpha, w_ce)
    assert relerr(dy, gref) < (1e-4 if dtype == torch.float32 else 8e-3)


def test_kd_token_loss_all_pad_and_identity():
    g = torch.Generator().manual_seed(3)
    y = torch.randn(3, 4, 64, generator=g)
    out5, dy, nval, _ = _token_loss(y, y.clone(), torch.zeros(3, 4, dtype=torch.long), 4.0, 1.0, 0.0, torch.float32)
    assert nval == 0 and abs(out5[2].item()) < 1e-6          # KL(p||p) = 0, no NaN in the gradient with zero valid rows
    assert torch.isfinite(dy).all() and dy.abs().max() < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", ["both", "feat_only", "hid_only", "truncated"])
def test_aux_loss_kernel(dtype, case):
    from imagecaptioner_b200.distillation_utils import DistillationLoss
    g = torch.Generator().manual_seed(11)
    B, S, E, T, H = 5, 49, 40, 6, 72
    fs = torch.randn(B, S, E, generator=g); ft = torch.randn(B, S, E, generator=g)
    hs = torch.randn(T, B, H, generator=g); ht = torch.randn(T if case != "truncated" else T - 2, B, H, generator=g)
    mod = DistillationLoss(alpha=0.0, beta=0.35, gamma=0.15, temperature=1.0, vocab_size=8)
    fs_d = fs.to(DEV).to(dtype).requires_grad_(True); ft_d = ft.to(DEV).requires_grad_(True)
    hs_d = hs.to(DEV).to(dtype).requires_grad_(True)
    logits = torch.zeros(T, B, 8, device=DEV)
    s_out, t_out = {"logits": logits}, {"logits": logits}
    if case != "hid_only":
        s_out["encoder_features"], t_out["encoder_features"] = fs_d, ft_d
    if case != "feat_only":
        s_out["hidden_states"], t_out["hidden_states"] = list(hs_d.unbind(0)), list(ht.to(DEV).unbind(0))
    loss, d = mod(s_out, t_out, torch.ones(T, B, dtype=torch.long, device=DEV))
    (loss * 3.0).backward()                                  # exercises the grad_output scaling
    fs_r = fs_d.detach().float().cpu().double().requires_grad_(True); ft_r = ft.double().requires_grad_(True)
    hs_r = hs_d.detach().float().cpu().double().requires_grad_(True)
    feat = O.feature_kd(fs_r, ft_r) if case != "hid_only" else torch.tensor(0.0)
    hid = O.hidden_kd(list(hs_r.unbind(0)), list(ht.double().unbind(0))) if case != "feat_only" else torch.tensor(0.0)
    ref = 0.35 * feat + 0.15 * hid
    (ref * 3.0).backward()
    assert abs(d["feature_kd_loss"] - float(feat)) < 1e-4 * abs(float(feat)) + 1e-7
    assert abs(d["hidden_kd_loss"] - float(hid)) < 1e-4 * abs(float(hid)) + 1e-7
    tol = 1e-4 if dtype == torch.float32 else 8e-3
    if case != "hid_only":
        assert relerr(fs_d.grad.float().cpu(), fs_r.grad) < tol and relerr(ft_d.grad.cpu(), ft_r.grad) < tol
    if case != "feat_only":
        assert relerr(hs_d.grad.float().cpu(), hs_r.grad) < tol


def test_loss_error_behaviour():
    """ValueError on feature / hidden width mismatch like the reference (distillation_utils.py:74,120)."""
    from imagecaptioner_b200.distillation_utils import DistillationLoss
    mod = DistillationLoss(vocab_size=8)
    lg = torch.zeros(2, 3, 8, device=DEV)
    tg = torch.ones(2, 3, dtype=torch.long, device=DEV)
    with pytest.raises(ValueError, match="Feature dimensions"):
        mod({"logits": lg, "encoder_features": torch.zeros(3, 49, 8, device=DEV)},
            {"logits": lg, "encoder_features": torch.zeros(3, 49, 16, device=DEV)}, tg)
    with pytest.raises(ValueError, match="Hidden dimensions"):
        mod({"logits": lg, "hidden_states": [torch.zeros(3, 8, device=DEV)] * 2},
            {"logits": lg, "hidden_states": [torch.zeros(3, 16, device=DEV)] * 2}, tg)
    # teacher hiddens None (what TeacherWrapper emits): hidden term is exactly 0
    _, d = mod({"logits": lg, "hidden_states": [torch.zeros(3, 8, device=DEV)] * 2}, {"logits": lg, "hidden_states": None}, tg)
    assert d["hidden_kd_loss"] == 0.0
    # the reference's single-term methods
    y = torch.randn(2, 3, 8, device=DEV); z = torch.randn(2, 3, 8, device=DEV)
    kd = mod.token_level_distillation(y, z)
    assert abs(float(kd) - float(O.token_kd(y.cpu(), z.cpu(), 4.0))) < 1e-5
    assert mod.decoder_hidden_state_distillation(None, None) == 0.0


def test_attention_step_accessor():
    from imagecaptioner_b200.student_model import LSTMDecoder
    torch.manual_seed(0)
    dec = LSTMDecoder(50, 32, 64, 2, dropout=0.0).to(DEV)
    h = torch.randn(5, 64, device=DEV); f = torch.randn(5, 49, 32, device=DEV)
    ctx, w = dec.attention_mechanism(h, f)
    rc, rw = O.attention_step(h.cpu(), f.cpu(), dec.attention.weight.detach().cpu(), dec.attention.bias.detach().cpu())
    assert relerr(ctx.cpu(), rc) < 1e-5 and relerr(w.cpu(), rw) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims", [(6, 49, 256, 4), (3, 49, 48, 4), (2, 20, 64, 4)])
def test_refinement_block(dims, dtype):
    """AttentionRefinement through b2c_refinement_forward / _backward vs the oracle's restatement under autograd."""
    from imagecaptioner_b200.student_model import AttentionRefinement
    B, S, E, heads = dims
    params = {k: v for k, v in O.init_student_params(50, E, 2 * E, 1, True, seed=12).items() if k.startswith("attention_refinement.")}
    for k in params:                                   # non-trivial biases / LayerNorm affine
        if "bias" in k or "norm" in k:
            params[k] = params[k] + 0.1 * torch.randn(params[k].shape, generator=torch.Generator().manual_seed(len(k)))
    mod = AttentionRefinement(E, heads).to(DEV).eval()
    mod.load_state_dict({k[len("attention_refinement."):]: v for k, v in params.items()})
    mod.compute_dtype = dtype
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, S, E, generator=g); w = torch.randn(B, S, E, generator=g)
    xd = x.to(DEV).requires_grad_(True)
    out = mod(xd)
    (out.float() * w.to(DEV)).sum().backward()
    P = {k: v.double().requires_grad_(True) for k, v in params.items()}
    xr = x.double().requires_grad_(True)
    ref = O.refinement_forward(P, xr, heads)
    (ref * w.double()).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 3e-2       # bf16 at these tiny widths; the 2e-2 bar is asserted at config-1 shapes
    assert relerr(out.float().cpu(), ref.detach()) < tol
    assert relerr(xd.grad.cpu(), xr.grad) < tol
    for k, v in mod.named_parameters():
        # bf16: the FFN's first Linear sits directly behind the ReLU; with this test's random output weighting its gradient is
        # a random-sign sum, so the ~0.3 % of mask elements that flip under bf16 forward rounding show up at full size
        # (DESIGN.md "bf16 parity").  fp32 holds 1e-4 on everything.
        lim = tol * (8 if (dtype == torch.bfloat16 and k.startswith("ffn.0")) else 1)
        assert relerr(v.grad.cpu(), P["attention_refinement." + k].grad) < lim, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims", [(4, 197, 384, 256, 49), (2, 197, 384, 256, 64), (3, 197, 40, 32, 49), (2, 49, 64, 32, 49), (2, 197, 64, 64, 49)])
def test_feature_projector(dims, dtype):
    """FeatureProjector through b2c_projector_forward / _backward vs the oracle; includes the reference's own shape pin
    (384 -> 256, 197 -> 64 gives (2,64,256), test_dimension_fix.py:16-43), no pooling (49 -> 49) and the identity projection."""
    from imagecaptioner_b200.distillation_utils import FeatureProjector
    B, St, Et, Es, So = dims
    pparams = O.init_projector_params(Et, Es, seed=4)
    for k in pparams:
        if k.endswith("bias") or ".3." in k:
            pparams[k] = pparams[k] + 0.1 * torch.randn(pparams[k].shape, generator=torch.Generator().manual_seed(len(k)))
    mod = FeatureProjector(Et, Es, St, So).to(DEV).eval()
    if pparams:
        mod.load_state_dict(pparams)
    mod.compute_dtype = dtype
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, St, Et, generator=g); w = torch.randn(B, So, Es, generator=g)
    out = mod(x.to(DEV))
    assert tuple(out.shape) == (B, So, Es) and out.dtype == torch.float32
    P = {k: v.double().requires_grad_(True) for k, v in pparams.items()}
    ref = O.feature_projector(P, x.double(), So)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert relerr(out.cpu(), ref.detach()) < tol
    if pparams:
        (out * w.to(DEV)).sum().backward()
        (ref * w.double()).sum().backward()
        for k, v in mod.named_parameters():
            lim = tol * (8 if (dtype == torch.bfloat16 and ".0." in k) else 1)      # ReLU-gated Linear, see test_refinement_block
            assert relerr(v.grad.cpu(), P[k].grad) < lim, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decoder_prepare_split_is_bit_identical(dtype):
    """b2c_decoder_prepare (side stream) + b2c_decoder_forward_prepared == b2c_decoder_forward, forward and backward."""
    from imagecaptioner_b200 import _ops
    V, E, H, L, B, T, S = 120, 64, 128, 2, 8, 5, 9
    params = O.init_student_params(V, E, H, L, False, seed=4)
    plist = [params["decoder." + k].to(DEV).requires_grad_(True) for k in _ops.param_order(L)]
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, S, E, generator=g).to(DEV)
    cap = torch.randint(1, V, (T, B), generator=g).to(DEV)
    w = torch.randn(T, B, V, generator=g).to(DEV)
    outs = []
    for split in (False, True):
        for p in plist:
            p.grad = None
        f = feats.clone().requires_grad_(True)
        prepared = _ops.decoder_prepare(cap, S, dtype, L, plist) if split else None
        logits, hid, attw = _ops.DecoderFunction.apply(f, cap, dtype, 0.0, 0, L, prepared, None, None, *plist)
        (logits.float() * w).sum().backward()
        torch.cuda.synchronize()
        outs.append([logits.detach().clone(), hid.detach().clone(), attw.clone(), f.grad.clone()] + [p.grad.clone() for p in plist])
    for i, (a, b) in enumerate(zip(*outs)):
        if i < 3:
            assert torch.equal(a, b)                       # forward: the same kernels on the same operands
        else:
            assert relerr(a, b) < 1e-6                     # backward: split-K reductions may add in a different order
    with pytest.raises(ValueError, match="prepared for"):
        _ops.DecoderFunction.apply(feats[:4], cap[:, :4], dtype, 0.0, 0, L, _ops.decoder_prepare(cap, S, dtype, L, plist), None, None, *plist)


def test_bleu1_kernel_matches_reference_fixture():
    import os
    from imagecaptioner_b200 import _ops
    from tests.harness import GOLDEN
    fx = torch.load(os.path.join(GOLDEN, "validation_case.pt"), weights_only=False)["bleu"]
    got = _ops.bleu1(fx["pred"].to(DEV), fx["targets"].to(DEV)).cpu()
    assert torch.allclose(got, fx["reference"], atol=1e-6)
    g = torch.Generator().manual_seed(2)                       # longer captions than a warp, many samples
    pred = torch.randint(0, 30, (70, 37), generator=g); tgt = torch.randint(0, 30, (70, 37), generator=g)
    assert torch.allclose(_ops.bleu1(pred.to(DEV), tgt.to(DEV)).cpu(), O.bleu1(pred, tgt), atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("V", [5000, 203])
def test_kd_token_eval_matches_training_kernel_and_argmax(V, dtype):
    """b2c_kd_token_eval: the same loss parts as the training pass (no gradient written) and logits.argmax(-1), lowest index on ties."""
    from imagecaptioner_b200 import _ops
    T, B = 3, 7
    g = torch.Generator().manual_seed(V)
    y = (torch.randn(T, B, V, generator=g) * 2).to(dtype)
    y[0, 0, 17] = y[0, 0].max() + 1; y[0, 0, 5] = y[0, 0, 17]            # a tie: the lower index wins
    z = torch.randn(T, B, V, generator=g) * 2
    tgt = torch.randint(0, V, (T, B), generator=g); tgt[1, :3] = 0
    yd, zd, td = y.to(DEV), z.to(DEV), tgt.to(DEV)
    out5, pred = _ops.kd_eval(yd, zd, td, None, None, None, None, 0.7, 0.0, 0.0, 4.0, 0.3)
    cfg = (0.7, 0.0, 0.0, 4.0, 0.3, 1.0, None, None, False)
    loss, ref5 = _ops.KDLossFunction.apply(yd.clone().requires_grad_(True), zd, td, None, None, None, None, cfg)
    assert torch.allclose(out5.cpu(), ref5.cpu(), rtol=2e-5, atol=1e-6)
    assert torch.equal(pred.cpu().long(), y.float().argmax(dim=-1))
    assert int(pred[0, 0]) == 5
