"""CPU: pin oracle/optim_oracle.py (and the host-side schedule in imagecaptioner_b200/optim.py) against the torch library
calls the reference's training loop makes (src/train_student_kd.py:230-236, :290-303)."""
import math

import pytest
import torch

from oracle import optim_oracle as OO


def _make_groups(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [[(7, 5), (5,)], [(33, 9), (9,), (3, 3, 2)], [(16,), (4, 6)]]
    return [[torch.randn(*s, generator=g, dtype=torch.float64) for s in grp] for grp in shapes]


def _rand_grads(groups, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return [[torch.randn(p.shape, generator=g, dtype=torch.float64) * scale for p in grp] for grp in groups]


@pytest.mark.parametrize("grad_scale", [0.01, 3.0])          # clip inactive / active
def test_oracle_matches_torch_adamw_and_two_clip_groups(grad_scale):
    vals = _make_groups()
    lrs = [1e-4, 1e-3, 1e-3]
    tparams = [[torch.nn.Parameter(p.clone()) for p in grp] for grp in vals]
    opt = torch.optim.AdamW([{"params": grp, "lr": lr} for grp, lr in zip(tparams, lrs)], weight_decay=0.01)
    # reference: groups 0+1 are "student_model.parameters()" (one clip call), group 2 is a projector (its own clip call)
    oracle = OO.OptimizerOracle([{"params": grp, "lr": lr, "clip_group": cg} for grp, lr, cg in zip(vals, lrs, [0, 0, 1])], max_norm=1.0)
    for it in range(5):
        grads = _rand_grads(vals, 100 + it, grad_scale)
        for grp, gg in zip(tparams, grads):
            for p, g in zip(grp, gg):
                p.grad = g.clone()
        n0 = torch.nn.utils.clip_grad_norm_(tparams[0] + tparams[1], max_norm=1.0)
        n1 = torch.nn.utils.clip_grad_norm_(tparams[2], max_norm=1.0)
        opt.step()
        oracle.step(grads)
        assert abs(oracle.last_norms[0] - float(n0)) < 1e-12 * max(1.0, float(n0))
        assert abs(oracle.last_norms[1] - float(n1)) < 1e-12 * max(1.0, float(n1))
        for grp, ogrp in zip(tparams, oracle.groups):
            for p, q in zip(grp, ogrp["params"]):
                assert torch.allclose(p.detach(), q, rtol=1e-12, atol=1e-14)


def test_oracle_grad_scaler_semantics():
    """unscale_, skipped step on inf/nan, backoff and growth of the scale: against torch.amp.GradScaler's documented rules,
    driven exactly as the reference drives it (unscale_ -> clip -> step -> update)."""
    vals = _make_groups(1)
    oracle = OO.OptimizerOracle([{"params": grp, "lr": 1e-3} for grp in vals], loss_scale=1024.0, growth_interval=3)
    ref = OO.OptimizerOracle([{"params": grp, "lr": 1e-3} for grp in vals], loss_scale=None)
    scale, tracker = 1024.0, 0
    for it in range(8):
        grads = _rand_grads(vals, 200 + it)
        scaled = [[g * scale for g in gg] for gg in grads]
        poisoned = it in (2, 6)
        if poisoned:
            scaled[1][0][0, 0] = float("inf") if it == 2 else float("nan")
        before = [p.clone() for g in oracle.groups for p in g["params"]]
        oracle.step(scaled)
        if poisoned:
            assert oracle.last_skipped
            assert all(torch.equal(a, b) for a, b in zip(before, [p for g in oracle.groups for p in g["params"]]))
            scale *= 0.5; tracker = 0
        else:
            ref.step(grads)                                    # the un-scaled optimizer takes only the clean steps
            tracker += 1
            if tracker == 3:
                scale *= 2.0; tracker = 0
        assert oracle.loss_scale == scale and oracle.growth_tracker == tracker
    assert oracle.step_count == ref.step_count == 6
    for ga, gb in zip(oracle.groups, ref.groups):
        for p, q in zip(ga["params"], gb["params"]):
            assert torch.allclose(p, q, rtol=1e-10, atol=1e-13)


def test_grad_scaler_library_agrees_on_cpu():
    """Same sequence through the actual torch.amp.GradScaler (CPU device) where this torch build supports it."""
    try:
        scaler = torch.amp.GradScaler("cpu", init_scale=256.0, growth_interval=2)
    except Exception as e:                                     # pragma: no cover
        pytest.skip(f"GradScaler('cpu') unavailable: {e}")
    vals = [[torch.randn(6, 4, dtype=torch.float32, generator=torch.Generator().manual_seed(3))]]
    tp = [torch.nn.Parameter(vals[0][0].clone())]
    opt = torch.optim.AdamW(tp, lr=1e-2, weight_decay=0.01)
    oracle = OO.OptimizerOracle([{"params": vals[0], "lr": 1e-2}], loss_scale=256.0, growth_interval=2, dtype=torch.float64)
    for it in range(6):
        g = torch.randn(6, 4, generator=torch.Generator().manual_seed(50 + it))
        if it == 3:
            g[0, 0] = float("inf")
        s = scaler.get_scale() if it else 256.0
        scaler.scale((tp[0] * g).sum()).backward()             # d/dp = g * scale, as scaler.scale(loss).backward() (:288)
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(tp, max_norm=1.0)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        oracle.step([[g * s]])
        assert scaler.get_scale() == oracle.loss_scale
        assert torch.allclose(tp[0].detach().double(), oracle.groups[0]["params"][0], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("T_0,T_mult", [(5, 2), (3, 1), (4, 3)])
def test_cosine_warm_restarts_matches_torch(T_0, T_mult):
    from imagecaptioner_b200.optim import cosine_warm_restarts
    base = [1e-5, 1e-4]
    ps = [torch.nn.Parameter(torch.zeros(1)) for _ in base]
    opt = torch.optim.SGD([{"params": [p], "lr": lr} for p, lr in zip(ps, base)])
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=T_0, T_mult=T_mult, eta_min=1e-6)
    n_batches = 7
    for epoch in range(40):
        for i in range(n_batches):
            e = epoch + i / n_batches                          # the reference's call: scheduler.step(epoch + batch_idx / len(loader))
            sched.step(e)
            for lr_t, b in zip(sched.get_last_lr(), base):
                assert math.isclose(OO.cosine_warm_restarts_lr(e, b, T_0, T_mult, 1e-6), lr_t, rel_tol=1e-9, abs_tol=1e-15), (e, T_0, T_mult)
                assert math.isclose(cosine_warm_restarts(e, b, T_0, T_mult, 1e-6), lr_t, rel_tol=1e-9, abs_tol=1e-15), (e, T_0, T_mult)


def test_flat_adamw_refuses_cpu_parameters():
    from imagecaptioner_b200.optim import FlatAdamW
    with pytest.raises(RuntimeError, match="GPU only"):
        FlatAdamW([torch.nn.Parameter(torch.zeros(4))])
