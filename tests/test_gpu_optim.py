"""GPU: the native optimizer step (b2c_optimizer_step via optim.FlatAdamW) against oracle/optim_oracle.py (fp64) and against the
torch library calls the reference makes (src/train_student_kd.py:230-236, :290-303)."""
import pytest
import torch

from oracle import optim_oracle as OO

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _groups(seed=0, big=False):
    g = torch.Generator().manual_seed(seed)
    shapes = [[(37, 5), (5,)], [(129, 33), (33,), (3, 7, 2)], [(50,), (64, 64)], [(19,), (7, 3)]]
    if big:
        shapes[1].append((1500, 1021))                       # > one pass of the grid, ragged tail
    return [[torch.randn(*s, generator=g) for s in grp] for grp in shapes]


def _grads(vals, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return [[torch.randn(p.shape, generator=g) * scale for p in grp] for grp in vals]


def _build(vals, lrs, clips, **kw):
    from imagecaptioner_b200.optim import FlatAdamW
    params = [[torch.nn.Parameter(p.clone().to(DEV)) for p in grp] for grp in vals]
    opt = FlatAdamW([{"params": grp, "lr": lr, "clip_group": cg} for grp, lr, cg in zip(params, lrs, clips)], **kw)
    return params, opt


def _set_grads(opt, params, grads):
    opt.zero_grad()
    for grp, gg in zip(params, grads):
        for p, g in zip(grp, gg):
            p.grad.copy_(g.to(DEV))                          # .grad is a view of the flat buffer


@pytest.mark.parametrize("grad_scale,big", [(0.01, False), (3.0, False), (1.0, True)])
def test_flat_adamw_matches_oracle(grad_scale, big):
    vals = _groups(0, big)
    lrs, clips = [1e-4, 1e-3, 1e-3, 1e-3], [0, 0, 0, 1]     # the reference: encoder / decoder / refinement share one clip call, projector its own
    params, opt = _build(vals, lrs, clips, max_grad_norm=1.0)
    oracle = OO.OptimizerOracle([{"params": grp, "lr": lr, "clip_group": cg} for grp, lr, cg in zip(vals, lrs, clips)], max_norm=1.0)
    assert all(p.data_ptr() >= opt.flat_param.data_ptr() for grp in params for p in grp)          # parameters live in the flat buffer
    for it in range(4):
        grads = _grads(vals, 10 + it, grad_scale)
        _set_grads(opt, params, grads)
        opt.step()
        oracle.step(grads)
        norms = opt.grad_norms()
        assert abs(norms[0] - oracle.last_norms[0]) <= 2e-6 * oracle.last_norms[0]
        assert abs(norms[1] - oracle.last_norms[1]) <= 2e-6 * oracle.last_norms[1]
        assert not opt.last_step_skipped()
    assert int(opt.step_count.item()) == 4
    for grp, ogrp in zip(params, oracle.groups):
        for p, q in zip(grp, ogrp["params"]):
            assert torch.allclose(p.detach().cpu().double(), q, rtol=2e-6, atol=2e-7)
    # moments too (they decide every later step)
    o = 0
    for g, ogrp in zip(opt.param_groups, oracle.groups):
        o = g["range"][0]
        for q_m, q_v in zip(ogrp["m"], ogrp["v"]):
            n = q_m.numel()
            assert torch.allclose(opt.exp_avg[o:o + n].cpu().double(), q_m.flatten(), rtol=5e-6, atol=1e-9)
            assert torch.allclose(opt.exp_avg_sq[o:o + n].cpu().double(), q_v.flatten(), rtol=5e-6, atol=1e-12)
            o += n


def test_flat_adamw_matches_torch_library_calls():
    """Same sequence through torch.optim.AdamW + two clip_grad_norm_ calls on the GPU (what the reference executes)."""
    vals = _groups(1)
    lrs, clips = [1e-4, 1e-3, 1e-3, 1e-3], [0, 0, 0, 1]
    params, opt = _build(vals, lrs, clips, max_grad_norm=1.0)
    tparams = [[torch.nn.Parameter(p.clone().to(DEV)) for p in grp] for grp in vals]
    topt = torch.optim.AdamW([{"params": grp, "lr": lr} for grp, lr in zip(tparams, lrs)], weight_decay=0.01)
    for it in range(3):
        grads = _grads(vals, 30 + it, 2.0)
        _set_grads(opt, params, grads)
        for grp, gg in zip(tparams, grads):
            for p, g in zip(grp, gg):
                p.grad = g.to(DEV).clone()
        torch.nn.utils.clip_grad_norm_(tparams[0] + tparams[1] + tparams[2], max_norm=1.0)
        torch.nn.utils.clip_grad_norm_(tparams[3], max_norm=1.0)
        topt.step()
        opt.step()
    for grp, tgrp in zip(params, tparams):
        for p, q in zip(grp, tgrp):
            assert torch.allclose(p.detach(), q.detach(), rtol=3e-6, atol=3e-7)


def test_loss_scale_skip_backoff_growth_and_lr_schedule():
    from imagecaptioner_b200.optim import CosineWarmRestarts
    vals = _groups(2)
    lrs, clips = [1e-4, 1e-3, 1e-3, 1e-3], [0, 0, 0, 1]
    params, opt = _build(vals, lrs, clips, max_grad_norm=1.0, loss_scale=1024.0, growth_interval=3)
    sched = CosineWarmRestarts(opt, T_0=5, T_mult=2, eta_min=1e-6)
    oracle = OO.OptimizerOracle([{"params": grp, "lr": lr, "clip_group": cg} for grp, lr, cg in zip(vals, lrs, clips)], max_norm=1.0,
                                loss_scale=1024.0, growth_interval=3)
    for it in range(9):
        epoch = it / 3.0                                        # scheduler.step(epoch + batch_idx / len(loader))
        sched.step(epoch)
        for g, b in zip(oracle.groups, lrs):
            g["lr"] = OO.cosine_warm_restarts_lr(epoch, b, 5, 2, 1e-6)
        scale = oracle.loss_scale
        grads = [[g * scale for g in gg] for gg in _grads(vals, 60 + it)]
        if it in (2, 7):
            grads[1][0][5, 5] = float("nan") if it == 2 else float("-inf")
        before = opt.flat_param.clone()
        _set_grads(opt, params, grads)
        opt.step()
        oracle.step(grads)
        assert opt.last_step_skipped() == oracle.last_skipped == (it in (2, 7))
        if oracle.last_skipped:
            assert torch.equal(before, opt.flat_param)          # parameters untouched by a skipped step
        assert float(opt.loss_scale.item()) == oracle.loss_scale
        assert int(opt.growth_tracker.item()) == oracle.growth_tracker
        assert int(opt.step_count.item()) == oracle.step_count
    for grp, ogrp in zip(params, oracle.groups):
        for p, q in zip(grp, ogrp["params"]):
            assert torch.allclose(p.detach().cpu().double(), q, rtol=3e-6, atol=3e-7)


def test_folded_data_parallel_average():
    """fold_average(world): gradients that are SUMS over `world` ranks give the same update (and norms) as the averaged ones."""
    vals = _groups(3)
    lrs, clips = [1e-4, 1e-3, 1e-3, 1e-3], [0, 0, 0, 1]
    params_a, opt_a = _build(vals, lrs, clips, max_grad_norm=1.0)
    params_b, opt_b = _build(vals, lrs, clips, max_grad_norm=1.0)
    opt_b.fold_average(4)
    for it in range(3):
        grads = _grads(vals, 80 + it, 1.5)
        _set_grads(opt_a, params_a, grads)
        _set_grads(opt_b, params_b, [[g * 4.0 for g in gg] for gg in grads])
        opt_a.step(); opt_b.step()
        na, nb = opt_a.grad_norms(), opt_b.grad_norms()
        assert abs(na[0] - nb[0]) <= 1e-6 * na[0] and abs(na[1] - nb[1]) <= 1e-6 * na[1]
    assert torch.allclose(opt_a.flat_param, opt_b.flat_param, rtol=1e-6, atol=1e-8)


def test_optimizer_step_rejects_bad_segments():
    import ctypes
    from imagecaptioner_b200 import _ops
    lib = _ops.load_library()
    buf = torch.zeros(64, device=DEV)
    st = torch.zeros(_ops.B2C_OPT_NSTATS, device=DEV); scratch = torch.zeros(_ops.B2C_OPT_SCRATCH_BYTES, dtype=torch.uint8, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV); lr = torch.ones(1, device=DEV)
    hp = _ops.B2COptHyper(0.9, 0.999, 1e-8, 1.0, 1.0, 2.0, 0.5, 2000)
    def call(seg):
        segs = (_ops.B2COptSegment * 1)(seg)
        return lib.b2c_optimizer_step(buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), segs, 1, ctypes.byref(hp), lr.data_ptr(), 1,
                                      step.data_ptr(), None, None, st.data_ptr(), scratch.data_ptr(), 0)
    assert call(_ops.B2COptSegment(2, 10, 0, 0, 0.01)) == -1            # begin not a multiple of 4
    assert b"segment" in lib.b2c_last_error()
    assert call(_ops.B2COptSegment(0, 10, 3, 0, 0.01)) == -1            # lr_index outside lr[]
    assert call(_ops.B2COptSegment(0, 10, 0, 9, 0.01)) == -1            # clip group out of range
    assert call(_ops.B2COptSegment(0, 10, 0, 0, 0.01)) == 0
    torch.cuda.synchronize()


def test_graphed_step_with_native_optimizer_matches_torch_optimizer():
    """GraphedKDStep with FlatAdamW (one clip group here, so both sides clip the same global norm) == the torch-optimizer step."""
    from imagecaptioner_b200.ddp import FlatGradAllReducer
    from imagecaptioner_b200.distillation_utils import DistillationLoss
    from imagecaptioner_b200.graph import GraphedKDStep
    from imagecaptioner_b200.optim import FlatAdamW
    from tests.harness import build_student
    from oracle import kd_oracle as O
    V, E, H, L, B, T = 200, 64, 128, 2, 8, 6
    params = O.init_student_params(V, E, H, L, True, seed=0)
    pparams = O.init_projector_params(48, E, seed=1)
    batch = {k: (v.to(DEV) if v is not None else None) for k, v in O.synthetic_batch(B, T, V, E, H, Et=48, seed=7).items()}
    results = []
    for native in (False, True):
        model, projector = build_student(params, pparams, V, E, H, L, True, 48, DEV)
        model.decoder.compute_dtype = torch.float32
        trainable = [p for p in list(model.parameters()) + list(projector.parameters()) if p.requires_grad]
        if native:
            opt = FlatAdamW(trainable, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
            reducer = None
        else:
            reducer = FlatGradAllReducer(trainable)
            opt = torch.optim.AdamW(trainable, lr=1e-3, weight_decay=0.01, fused=True, capturable=True)
        kd = GraphedKDStep(model, projector, DistillationLoss(vocab_size=V), opt, reducer, batch, autocast_dtype=None, use_graph=False)
        outs = [kd.step().clone() for _ in range(3)]
        torch.cuda.synchronize()
        results.append((torch.stack(outs).cpu(), {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}))
    (l0, w0), (l1, w1) = results
    assert torch.allclose(l0, l1, rtol=2e-4, atol=1e-6)
    for k in w0:
        assert torch.allclose(w0[k], w1[k], rtol=1e-3, atol=2e-5), k
