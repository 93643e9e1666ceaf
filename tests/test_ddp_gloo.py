"""CPU, world_size 2 over gloo: the host-side data-parallel logic (flat gradient buffer all-reduce, global non-PAD
count) reproduces single-process global-batch gradients.  The loss arithmetic here is the ORACLE's (no GPU in this
container); the GPU form of the same check is tests/test_gpu_parity.py::test_data_parallel_normaliser_matches_global_batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kd_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_reference(params, batch):
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out, hids, _ = O.decoder_forward(P, batch["encoder_features"], batch["captions_input"])
    loss = 0.5 * O.cross_entropy_ignore_pad(out, batch["targets"]) + 0.5 * O.token_kd(out, batch["teacher_logits"], 4.0)
    loss.backward()
    return float(loss), {k: v.grad.clone() for k, v in P.items()}


def _worker(rank, world, port, params, batch, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from imagecaptioner_b200.ddp import FlatGradAllReducer, attach_loss_group, shard_batch
        from imagecaptioner_b200.distillation_utils import DistillationLoss
        sl = shard_batch(batch["targets"].shape[1], rank, world)
        plist = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in params.values()])
        P = dict(zip(params.keys(), plist))
        reducer = FlatGradAllReducer(plist)
        loss_mod = DistillationLoss(vocab_size=batch["teacher_logits"].shape[-1])
        attach_loss_group(loss_mod)
        assert loss_mod.world_size == world and loss_mod.process_group is not None
        out, _, _ = O.decoder_forward(P, batch["encoder_features"][sl], batch["captions_input"][:, sl])
        tgt = batch["targets"][:, sl]
        # what the CUDA loss does under DP: all-reduce the non-PAD count, scale CE by world / global count
        nval = torch.tensor([int((tgt != 0).sum())], dtype=torch.int32)
        dist.all_reduce(nval, group=loss_mod.process_group)
        y = out.reshape(-1, out.shape[-1])
        lse = torch.logsumexp(y, 1); picked = y.gather(1, tgt.reshape(-1, 1)).squeeze(1)
        ce = ((lse - picked) * (tgt.reshape(-1) != 0)).sum() * loss_mod.world_size / nval.item()
        loss = 0.5 * ce + 0.5 * O.token_kd(out, batch["teacher_logits"][:, sl], 4.0)
        reducer.zero_grad()
        loss.backward()
        assert all(p.grad.data_ptr() >= reducer.flat.data_ptr() for p in plist)        # grads are views of the flat buffer
        reducer.allreduce()
        lt = torch.tensor([float(loss)]); dist.all_reduce(lt)
        if rank == 0:
            ret["loss"] = float(lt) / world
            ret["grads"] = {k: p.grad.clone() for k, p in P.items()}
            ret["nval"] = int(nval)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gradients_equal_global_batch():
    V, E, H, L, B, T = 40, 16, 24, 2, 6, 4
    params = O.init_student_params(V, E, H, L, False, seed=2)
    batch = O.synthetic_batch(B, T, V, E, H, seed=3)
    batch["targets"][2:, 0] = 0                       # uneven PAD counts between the two shards
    ref_loss, ref_grads = _global_reference(params, batch)
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, params, batch, ret)) for r in range(2)]
        [p.start() for p in procs]
        [p.join(150) for p in procs]
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert ret["nval"] == int((batch["targets"] != 0).sum())
        assert abs(ret["loss"] - ref_loss) < 1e-5 * abs(ref_loss)
        for k, g in ref_grads.items():
            got = ret["grads"][k]
            assert float((got - g).abs().max()) <= 1e-5 * float(g.abs().max()) + 1e-9, k


def test_flat_buffer_and_sharding_single_process():
    from imagecaptioner_b200.ddp import FlatGradAllReducer, shard_batch
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2), requires_grad=False)]
    r = FlatGradAllReducer(ps)
    assert r.flat.numel() == 17 and r.world_size == 1
    (ps[0].sum() * 2 + ps[1].sum()).backward()
    assert torch.equal(r.flat, torch.cat([torch.full((12,), 2.0), torch.ones(5)]))
    assert r.allreduce() is None
    r.zero_grad()
    assert float(r.flat.abs().sum()) == 0 and ps[0].grad.data_ptr() == r.flat.data_ptr()
    assert shard_batch(4096, 3, 8) == slice(1536, 2048)
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)


def test_early_allreduce_segment_and_reference_param_groups():
    """Host logic of the overlapped gradient exchange: the decoder's parameters must form ONE contiguous range of the flat buffer
    (that range is all-reduced under the refinement backward); and the reference's LR / clip grouping."""
    import torch.nn as nn
    from imagecaptioner_b200.ddp import FlatGradAllReducer
    from imagecaptioner_b200.graph import GraphedKDStep
    from imagecaptioner_b200.optim import reference_param_groups

    class Tiny(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = nn.Linear(3, 3)
            self.attention_refinement = nn.Linear(4, 4)
            self.decoder = nn.Sequential(nn.Linear(5, 5), nn.Linear(5, 2))
            self.use_attention_refinement = True

    model, projector = Tiny(), nn.Linear(6, 6)
    for p in model.encoder.parameters():
        p.requires_grad = False                                   # frozen encoder, as in the bench (features are fed directly)
    groups = reference_param_groups(model, {"encoder": projector}, 1e-3)
    assert [len(g["params"]) for g in groups] == [2, 4, 2, 2]
    assert [g["lr"] for g in groups] == [1e-4, 1e-3, 1e-3, 1e-3] and [g["clip_group"] for g in groups] == [0, 0, 0, 1]
    assert groups[3]["lr_group"] == 2                             # projector shares the "other" learning-rate group

    def early(order):
        kd = GraphedKDStep.__new__(GraphedKDStep)
        kd.model = model
        kd.reducer = FlatGradAllReducer(order)
        return kd._early_segment()

    dec, ref, prj = list(model.decoder.parameters()), list(model.attention_refinement.parameters()), list(projector.parameters())
    n_dec = sum(p.numel() for p in dec)
    assert early(dec + ref + prj) == (0, n_dec)
    n_ref = sum(p.numel() for p in ref)
    assert early(ref + dec + prj) == (n_ref, n_ref + n_dec)
    assert early(dec[:2] + ref + dec[2:] + prj) is None           # a foreign parameter inside the decoder's range: no early exchange
