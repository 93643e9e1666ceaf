#!/usr/bin/env python
"""KD train-step throughput of the B200-native hot path (BASELINE.json metric: KD train samples/sec).

One "step" = one pass of the hot path over one synthetic batch, replaying the reference's training step
(src/train_student_kd.py:262-303) with the encoders outside the path: refinement + attention-LSTM decoder
forward, FeatureProjector, DistillationLoss (all four terms), backward to every decoder / refinement /
projector parameter and the encoder features, gradient all-reduce (N>1), global-norm clip and AdamW.

  python bench.py [--gpus N] [--steps K] [--warmup W]              # our arm  (torchrun launches N>1)
  python bench.py --impl reference ...                             # the reference algorithm on the host CPU cores

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same step through the
public modules with every step's inputs copied from pinned host memory and the loss read back.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json configs[1]: default student, batch 512 per GPU, len 20, vocab 5000, ViT-small teacher features 197x384
CFG = dict(B=512, T=20, V=5000, E=256, H=512, L=2, S=49, St=197, Et=384)
METRIC = "kd_train_samples_per_sec"


def log(msg):
    if os.environ.get("BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - T0:7.2f}s] {msg}", file=sys.stderr, flush=True)


T0 = time.perf_counter()


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(cfg, seed, dtype_teacher=torch.float32):
    from oracle import kd_oracle as O   # synthetic-input generator shared with the tests (data only, no compute)
    return O.synthetic_batch(cfg["B"], cfg["T"], cfg["V"], cfg["E"], cfg["H"], cfg["S"], cfg["St"], cfg["Et"], seed=seed)


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference algorithm (oracle port: plain tensor arithmetic + autograd on the host CPU) for the same step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import kd_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Bs = 16                                                   # bounded sample: BASELINE configs[0] batch (the reference loader's own cap)
    cfg = dict(CFG, B=Bs)
    params = O.init_student_params(cfg["V"], cfg["E"], cfg["H"], cfg["L"], True, seed=0)
    pparams = O.init_projector_params(cfg["Et"], cfg["E"], seed=1)
    batch = make_batch(cfg, 1234)
    leaves = {k: v.clone().requires_grad_(True) for k, v in {**params, **{"proj." + k: v for k, v in pparams.items()}}.items()}
    opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4, weight_decay=0.01)

    def step():
        P = {k: v for k, v in leaves.items() if not k.startswith("proj.")}
        Q = {k[5:]: v for k, v in leaves.items() if k.startswith("proj.")}
        feats = batch["encoder_features"].clone().requires_grad_(True)
        outputs, enc, hids, _ = O.student_forward(P, feats, batch["captions_input"], True)
        tproj = O.feature_projector(Q, batch["teacher_features"], cfg["S"])
        th = batch["teacher_hiddens"]
        total, _ = O.distillation_loss({"logits": outputs, "encoder_features": enc, "hidden_states": hids},
                                       {"logits": batch["teacher_logits"], "encoder_features": tproj,
                                        "hidden_states": [th[t] for t in range(th.shape[0])]}, batch["targets"])
        opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_(list(leaves.values()), 1.0)
        opt.step()
        return float(total.detach())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = Bs / dt
    sample = f"B={Bs} (configs[0] batch) of the configs[1] model, T=20 V=5000, fp32, {threads} threads, per-sample rate"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "BASELINE configs[1]: CaptioningStudent default (E256 H512 2-layer LSTM, 49x256 feats, refinement) KD train step, "
                        "batch 512/GPU, len 20, vocab 5000, synthetic ViT-small teacher logits + 197x384 features + hidden states, bf16 compute / fp32 master",
            "per_gpu_batch": CFG["B"], "global_batch": CFG["B"] * n, "seq_len": CFG["T"], "vocab": CFG["V"], "parallelism": f"dp{n}",
            "step": "refinement+decoder fwd, projector, 4-term KD loss, bwd, grad allreduce, clip, AdamW",
            "l2_policy": "inputs per step (~410 MB logits+teacher) exceed the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------- our arm (GPU)
def cpu_baseline(budget_s=20.0):
    """Oracle port on the host cores, bounded sample (rank 0, N=1 only)."""
    from oracle import kd_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = dict(CFG, B=16)
    params = O.init_student_params(cfg["V"], cfg["E"], cfg["H"], cfg["L"], True, seed=0)
    pparams = O.init_projector_params(cfg["Et"], cfg["E"], seed=1)
    batch = make_batch(cfg, 1234)
    O.kd_step(params, pparams, batch)                       # warm-up
    t0, n = time.perf_counter(), 0
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 40):
        O.kd_step(params, pparams, batch); n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": 16 / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{n} fwd+loss+bwd steps of B=16 (configs[0]) T=20 V=5000 fp32 with the oracle port, {threads} threads; per-sample rate"}


def run_ours(args):
    import torch.distributed as dist
    from imagecaptioner_b200 import _ops
    from imagecaptioner_b200.ddp import FlatGradAllReducer, attach_loss_group
    from imagecaptioner_b200.distillation_utils import DistillationLoss
    from imagecaptioner_b200.graph import GraphedKDStep
    from imagecaptioner_b200.optim import FlatAdamW, reference_param_groups
    from oracle import kd_oracle as O
    from tests.harness import build_student

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun for N>1"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    log(f"start world={world} local={local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        log("process group up")
    lib = _ops.load_library()
    cfg = CFG
    B, T, V, E, H, L = cfg["B"], cfg["T"], cfg["V"], cfg["E"], cfg["H"], cfg["L"]
    params = O.init_student_params(V, E, H, L, True, seed=0)
    pparams = O.init_projector_params(cfg["Et"], E, seed=1)
    model, projector = build_student(params, pparams, V, E, H, L, True, cfg["Et"], dev)   # eval(): dropout off (parity config)
    model.decoder.compute_dtype = torch.bfloat16
    loss_mod = DistillationLoss(0.7, 0.2, 0.1, 4.0, vocab_size=V)
    attach_loss_group(loss_mod)
    if args.torch_optimizer:                 # A/B: torch's fused AdamW + one global-norm clip instead of the native flat-buffer step
        trainable = [p for p in list(model.parameters()) + list(projector.parameters()) if p.requires_grad]
        reducer = FlatGradAllReducer(trainable)
        opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=0.01, fused=True, capturable=not args.no_graph)
    else:                                    # the reference's LR groups + two clip groups (train_student_kd.py:219-234, :293-297)
        opt = FlatAdamW(reference_param_groups(model, projector, 1e-4), weight_decay=0.01, max_grad_norm=1.0)
        reducer = opt.reducer

    host = make_batch(cfg, 1234 + rank)
    keys = list(GraphedKDStep.INPUT_KEYS)
    pinned = {k: host[k].pin_memory() for k in keys}
    resident = {k: pinned[k].to(dev, non_blocking=True) for k in keys}
    h2d_bytes = sum(pinned[k].numel() * pinned[k].element_size() for k in keys)
    loss_host = torch.zeros(5, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    # the whole step (fwd, loss, bwd, all-reduce, clip, AdamW) is ONE CUDA graph over static input buffers
    n_before = lib.b2c_launch_count()
    kd = GraphedKDStep(model, projector, loss_mod, opt, reducer, resident, max_grad_norm=1.0, autocast_dtype=torch.bfloat16,
                       use_graph=not args.no_graph, warmup_steps=3)
    log("step object ready (warm-up + capture done)")
    launches_per_step = None if args.no_graph else (lib.b2c_launch_count() - n_before) // 4      # 3 warm-up bodies + 1 captured body

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM (already in the static buffers)
    sampler = ClockSampler(local)            # clocks / throttle reasons under load: started with the warm-up replays, 20 ms period
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup if args.profile else max(args.warmup, 3)):
        kd.step()
    barrier_sync()
    n0 = lib.b2c_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync()
    ev0.record()
    for _ in range(args.steps):
        out5 = kd.step()
    ev1.record()
    barrier_sync()
    ms = ev0.elapsed_time(ev1)
    if launches_per_step is None:
        launches_per_step = (lib.b2c_launch_count() - n0) / args.steps
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = B * world / (ms_step * 1e-3)
    final_loss = out5.tolist()
    log(f"value leg done: {ms_step:.3f} ms/step")

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_step, "value": value, "gpu_launches_per_step": launches_per_step}), flush=True)
        finish(world)
        return
    # ---- e2e: every step's inputs come from pinned host memory (uploaded on a copy stream into a staging set while the
    # previous step computes, then moved device-to-device into the graph's static buffers); the loss goes back to the host
    copy_stream = torch.cuda.Stream()
    staging = {k: torch.empty_like(resident[k]) for k in keys}
    staged_evt, consumed_evt = torch.cuda.Event(), torch.cuda.Event()

    def upload():
        copy_stream.wait_event(consumed_evt)
        with torch.cuda.stream(copy_stream):
            for k in keys:
                staging[k].copy_(pinned[k], non_blocking=True)
            staged_evt.record(copy_stream)

    def e2e_loop(n):
        consumed_evt.record()
        upload()
        for i in range(n):
            torch.cuda.current_stream().wait_event(staged_evt)
            kd.load(staging)
            consumed_evt.record()
            if i + 1 < n:
                upload()                                               # next step's H2D overlaps this step's compute
            o5 = kd.step()
            loss_host.copy_(o5, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_loop(2)
    barrier_sync()
    ev0.record()
    e2e_loop(args.steps)
    ev1.record()
    barrier_sync()
    e_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_val = B * world / (float(e_ms.item()) / args.steps * 1e-3)
    log("e2e leg done")

    # ---- roofline of the named kernels, timed live with CUDA events on the launching stream
    peaks = read_peaks()
    roof, extra = None, {}
    if rank == 0:
        roof, extra = kernel_rooflines(lib, _ops, dev, cfg, peaks)
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3) + 3,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 20},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches_per_step, "cuda_graph": not args.no_graph,
            "optimizer": "torch fused AdamW + global clip" if args.torch_optimizer else "native b2c_optimizer_step (3 LR groups, 2 clip groups)",
            "roofline": roof, "kernels": extra, "loss": final_loss}
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline() if world == 1 and not args.no_cpu_baseline else None
        print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Multi-rank teardown: barrier, flush, hard exit.  destroy_process_group() after CUDA-graph work has been seen to
    block for minutes on this stack (round 1, 2 GPUs: the JSON line was printed, then the job hung until the box timeout)."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def run_decode(args):
    """BASELINE configs[3]: greedy caption decode, batch 2048, max length 30 (one device-side decode per step of this leg).
    Reports captions/s and tokens/s for the bf16 throughput mode and the fp32 token-id-parity mode, plus the oracle on the CPU."""
    from imagecaptioner_b200 import _ops
    from oracle import kd_oracle as O
    from tests.harness import build_student
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    V, E, H, L, B, max_len = CFG["V"], CFG["E"], CFG["H"], CFG["L"], 2048, 30
    params = O.init_student_params(V, E, H, L, False, seed=0, logit_scale=8.0)
    model, _ = build_student(params, {}, V, E, H, L, False, E, dev)
    feats = torch.randn(B, CFG["S"], E, device=dev)
    out = {"workload": "BASELINE configs[3]: greedy decode batch 2048, max len 30, default student", "steps": args.steps}
    for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        model.decoder.compute_dtype = dt
        for _ in range(3):
            model.decoder.greedy(feats, max_len)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(args.steps):
            toks, lens = model.decoder.greedy(feats, max_len)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {"ms_per_decode": ms, "captions_per_s": B / ms * 1e3, "tokens_per_s": B * max_len / ms * 1e3}
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Bc = 64
    fc = torch.randn(Bc, CFG["S"], E)
    t0 = time.perf_counter(); O.greedy_decode(params, fc, max_len); dt_cpu = time.perf_counter() - t0
    out["cpu_baseline"] = {"captions_per_s": Bc / dt_cpu, "tokens_per_s": Bc * max_len / dt_cpu, "cores": threads, "kind": "port", "sample": f"one batched oracle decode of B={Bc}"}
    print(json.dumps(out), flush=True)


def kernel_rooflines(lib, _ops, dev, cfg, peaks):
    """Per-kernel achieved bandwidth / FLOP rate: algorithmic bytes / CUDA-event time per launch."""
    B, T, V, E, H = cfg["B"], cfg["T"], cfg["V"], cfg["E"], cfg["H"]
    N = T * B
    st = torch.cuda.current_stream().cuda_stream
    y = torch.randn(N, V, device=dev).bfloat16()
    z = torch.randn(N, V, device=dev) * 2
    tgt = torch.randint(1, V, (N,), device=dev)
    nval = torch.tensor([N], dtype=torch.int32, device=dev)
    dy = torch.empty_like(y)
    rows = torch.empty(2, N, device=dev)

    def kd():
        rc = lib.b2c_kd_token_loss(y.data_ptr(), z.data_ptr(), tgt.data_ptr(), N, V, 4.0, 0.7, 0.0, 1.0, nval.data_ptr(), dy.data_ptr(),
                                   rows[0].data_ptr(), rows[1].data_ptr(), _ops.B2C_BF16, torch.cuda.current_stream().cuda_stream)
        assert rc == 0

    def timeit(fn, reps=20):
        """Average device time of one launch: `reps` back-to-back launches captured in a CUDA graph (so the host-side ctypes /
        tensor-map cost of a launch, ~10 us, is not what is measured), timed with CUDA events around 5 replays."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (5 * reps) * 1e-3

    t_kd = timeit(kd)
    kd_bytes = N * V * (2 + 4 + 2) + 8 * N                     # read bf16 student + fp32 teacher, write bf16 dlogits, read targets
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_kernel_traffic.json")       # dram__bytes_read + write of ONE ncu --set full capture
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get("kd_token_loss_pipe_kernel<bf16,3,true>")
        if rec:
            traffic, traffic_src = rec["dram_bytes_read"] + rec["dram_bytes_write"], rec["source"]
    roof = {"kernel": "kd_token_loss_pipe_kernel<bf16,3,true> (b2c_kd_token_loss)", "bound": "hbm", "achieved": kd_bytes / t_kd / 1e9, "peak": peaks["hbm"],
            "unit": "GB/s", "frac": kd_bytes / t_kd / 1e9 / peaks["hbm"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["src"] + " (MEASURED_PEAKS.json hbm_gbs)",
            "algorithmic_bytes_per_launch": kd_bytes, "us_per_launch": t_kd * 1e6}
    # vocabulary head GEMM (time-batched, tcgen05): logits = o1 W2^T, M = T*B, N = V, K = E
    A = torch.randn(N, E, device=dev).bfloat16(); W = torch.randn(V, E, device=dev).bfloat16(); C = torch.empty(N, V, device=dev, dtype=torch.bfloat16)
    t_g = timeit(lambda: _ops.gemm(A, W, N, V, E, C=C))
    fl = 2.0 * N * V * E
    # recurrent gate GEMM of one step (serial part): M = B, N = 4H, K = E+H
    A2 = torch.randn(B, E + H, device=dev).bfloat16(); W2 = torch.randn(4 * H, E + H, device=dev).bfloat16(); C2 = torch.empty(B, 4 * H, device=dev)
    t_g2 = timeit(lambda: _ops.gemm(A2, W2, B, 4 * H, E + H, C=C2))
    fl2 = 2.0 * B * 4 * H * (E + H)
    extra = {"vocab_head_gemm_tcgen05": {"bound": "tensor", "achieved": fl / t_g / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                                         "frac": fl / t_g / 1e12 / peaks["tf_burst"], "us_per_launch": t_g * 1e6, "shape": [N, V, E]},
             "lstm_gate_gemm_tcgen05": {"bound": "tensor", "achieved": fl2 / t_g2 / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                                        "frac": fl2 / t_g2 / 1e12 / peaks["tf_burst"], "us_per_launch": t_g2 * 1e6, "shape": [B, 4 * H, E + H]}}
    return roof, extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b2c", choices=["b2c", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="kd_step", choices=["kd_step", "decode"], help="kd_step = the contract's metric (default); decode = configs[3] greedy decode leg")
    ap.add_argument("--torch-optimizer", action="store_true", help="A/B: torch fused AdamW + one global clip instead of the native optimizer step")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--profile", action="store_true", help="only warm-up + timed steps (for ncu launch lists); prints a reduced line")
    args = ap.parse_args()
    if args.workload == "decode":
        run_decode(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
