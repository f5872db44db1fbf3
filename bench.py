#!/usr/bin/env python
"""Benchmark of the hot path: utterances/sec of one training step (zero_grad -> forward -> label-smoothed CE ->
backward -> single gradient all-reduce -> Adam) of the speech Transformer, BASELINE.json cfg2 per GPU:
4 layers / 8 heads / d=512 / d_ff=2048 / vgg_cnn, batch 32, T_src=800 x 161 mel, T_tgt=100, V=4364, dropout 0.1.

    python bench.py --gpus N --steps K --warmup W                 # ours  (torchrun launches N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W    # the reference algorithm on the host CPU

Prints ONE JSON line (see the task contract): `value` = whole-job utt/s with inputs resident in HBM, `e2e` = the same
step through the public API with pinned-host inputs copied in and the loss read back every step, `roofline` for the
dominant kernel group (CUDA-event timed on the launching stream inside the timed region), `cpu_baseline` = the oracle
port on the host cores (rank 0, N=1).  Data are synthetic, weights random-init (no network for corpora/checkpoints).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "cfg2"
METRIC = "utterances/sec (fwd+bwd) dim512/4L/8H T_src=800"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        rows = self.rows[self.first:] or self.rows[-3:]     # a very short timed region can fall between two samples
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------------- algorithmic work
def algorithmic_flops(name, a):
    """FLOPs of one C-ABI call from its raw arguments (formulae of SURVEY.md §8d)."""
    if name == "linear_fwd" or name == "linear_bwd_data" or name == "linear_bwd_weight":
        M, N, K = a[4], a[5], a[6]
        return 2.0 * M * N * K
    if name in ("conv3x3_fwd", "conv3x3_bwd_data", "conv3x3_bwd_weight"):
        B, T, F, Ci, Co = a[5], a[6], a[7], a[8], a[9]
        return 2.0 * 9 * B * T * F * Ci * Co
    if name == "conv3x3_c1_fwd" or name == "conv3x3_c1_bwd_weight":
        off = 4
        B, F, T, Co = a[off], a[off + 1], a[off + 2], a[off + 3]
        return 2.0 * 9 * B * T * F * Co
    if name == "sdpa_fwd":
        B, H, Tq, Tk, dk, dv = a[20:26]
        return 2.0 * B * H * Tq * Tk * (dk + dv)
    if name == "sdpa_bwd":
        B, H, Tq, Tk, dk, dv = a[25:31]
        return 2.0 * 2.0 * B * H * Tq * Tk * (dk + dv)      # four GEMMs (recompute of QK^T not counted)
    if name == "sdpa_mat_fwd":      # q,k,v + 9 strides + key_pad, dense, causal, out + 3 strides, probs, probs_drop, then dims
        B, H, Tq, Tk, dk, dv = a[21:27]
        return 2.0 * B * H * Tq * Tk * (dk + dv)
    if name == "sdpa_mat_bwd":      # dout,q,k,v + 12 strides + 6 pointers, then dims
        B, H, Tq, Tk, dk, dv = a[22:28]
        return 2.0 * 2.0 * B * H * Tq * Tk * (dk + dv)
    return 0.0


def summarize_profile(report, steps, peaks):
    groups = []
    for name, calls in report.items():
        ms = sum(t for _, t in calls) / steps
        fl = sum(algorithmic_flops(name, a) for a, _ in calls) / steps
        groups.append(dict(kernel=name, launches_per_step=len(calls) / steps, ms_per_step=ms, gflop_per_step=fl / 1e9))
    groups.sort(key=lambda g: -g["ms_per_step"])
    tot = sum(g["ms_per_step"] for g in groups)
    for g in groups:
        g["share"] = g["ms_per_step"] / tot if tot else 0.0
        g["tflops"] = g["gflop_per_step"] / g["ms_per_step"] if g["ms_per_step"] else 0.0
    return groups


# ---------------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """Threads for the CPU arm: the PHYSICAL cores this process may use (torch's own default; one thread per hardware
    thread -- 128 on the 64-core GPU box -- was measured 30x slower), also under torchrun, which exports OMP_NUM_THREADS=1,
    and never more than a cgroup CPU quota allows."""
    n = None
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
    except Exception:
        pass
    logical = len(os.sched_getaffinity(0))
    n = min(n or max(1, logical // 2), logical)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def cpu_oracle_throughput(batch, steps, warmup, threads=None):
    """Oracle port (oracle/asr_oracle.py: the reference's algorithm restated on torch CPU kernels) fwd+bwd utt/s."""
    from oracle import asr_oracle as O
    import b200asr
    torch.set_num_threads(threads or host_threads())
    c = b200asr.BASELINE_CONFIGS[WORKLOAD]["cfg"]
    ocfg = O.OracleConfig(num_layers=c.num_layers, num_heads=c.num_heads, dim_model=c.dim_model, dim_key=c.dim_key,
                          dim_value=c.dim_value, dim_inner=c.dim_inner, vocab=c.vocab, feat_extractor=c.feat_extractor,
                          tgt_max_len=c.tgt_max_len, freq=c.freq)
    P = O.init_params(ocfg, seed=123456)
    src, lens, tgt = O.synthetic_batch(ocfg, batch, b200asr.BASELINE_CONFIGS[WORKLOAD]["t_src"], seed=0, ragged=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward_backward(P, ocfg, src, lens, tgt, c.label_smoothing)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return batch / mean, mean, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    b = 8            # bounded sample: a quarter of the cfg2 batch per step (~3 s of CPU work), enough rows to keep all cores busy
    ups, mean, threads = cpu_oracle_throughput(b, args.steps, args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": ups, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": f"{b} utterances per step of the cfg2 shape (T_src=800, T_tgt=100)"},
           "cpu_baseline": {"value": ups, "unit": "utt/s", "cores": threads, "kind": "port",
                            "sample": f"oracle fwd+bwd on {b} utterances/step, {args.steps} steps"},
           "e2e": {"value": ups, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    import b200asr
    import importlib
    ops = importlib.import_module(b200asr.__name__ + ".ops")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = b200asr._lib.load(check_device=True)
    if args.precision:
        names = ["linear", "conv", "attn", "conv_wgrad", "attn_bwd"]
        b200asr.precision.set(**dict(zip(names, args.precision.split(","))))
    spec = b200asr.BASELINE_CONFIGS[WORKLOAD]
    cfg, B, T = spec["cfg"], args.batch or spec["batch"], spec["t_src"]
    torch.manual_seed(123456)
    model = b200asr.build_model(cfg).to(dev)
    model.train()
    b200asr.manual_seed(1234 + rank)
    dp = b200asr.DataParallelStep(model, model_size=cfg.dim_input, warmup=4000, k_lr=1.0, min_lr=1e-6,
                                  smoothing=cfg.label_smoothing)
    g = torch.Generator().manual_seed(rank)
    src_h = torch.randn(B, 1, cfg.freq, T, generator=g).pin_memory()
    tgt_h = torch.randint(3, cfg.vocab, (B, cfg.tgt_max_len - 1), generator=g).pin_memory()
    lens = torch.full((B,), T, dtype=torch.int32)
    lens_d = lens.to(dev)
    src_d, tgt_d = src_h.to(dev), tgt_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident timed region (no instrumentation)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi takes a few 100 ms to produce its first row: start it before the warm-up ...
    for _ in range(args.warmup):
        dp.step(src_d, lens_d, tgt_d)
    barrier()
    sampler.mark()               # ... and keep only the rows sampled from here on (timed + instrumented regions, GPU busy)
    launches0 = lib.b200asr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dp.step(src_d, lens_d, tgt_d)
    e1.record()
    barrier()
    launches = (lib.b200asr_launch_count() - launches0) / args.steps
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    # ---- same steps again with a CUDA-event pair around every C-ABI call (per-kernel-group durations for the roofline);
    #      kept out of `value` because the extra host work per call can starve a ~40 ms step
    prof_steps = min(args.steps, 5)
    ops.profiler.start()
    e0.record()
    for _ in range(prof_steps):
        dp.step(src_d, lens_d, tgt_d)
    e1.record()
    barrier()
    ops.profiler.stop()
    ms_prof = e0.elapsed_time(e1) / prof_steps
    clocks = sampler.stop() if rank == 0 else None
    report = ops.profiler.report()

    # ---- end-to-end: pinned host inputs copied in, loss read back, every step
    loss_h = torch.zeros(1).pin_memory()
    for _ in range(min(2, args.warmup)):
        dp.step(src_h.to(dev, non_blocking=True), lens_d, tgt_h.to(dev, non_blocking=True))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        s = src_h.to(dev, non_blocking=True)
        t = tgt_h.to(dev, non_blocking=True)
        dp.step(s, lens_d, t)
        loss_h.copy_(dp.global_loss().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()         # the caller reads the loss every step
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    ms_e2e = max(max_over_ranks(e0.elapsed_time(e1)) / args.steps, 0.0)
    final_loss = float(loss_h.item())

    if rank != 0:
        return
    peaks = load_peaks()
    groups = summarize_profile(report, prof_steps, peaks)
    top = next((g for g in groups if g["gflop_per_step"] > 0), groups[0])
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0       # kind::tf32 runs at half the bf16 rate
    # DRAM bytes per launch (dram__bytes_read + dram__bytes_write) from the ncu --set full captures of this code under
    # profiles/ (engine_WgradPolicy_r1b.md, conv_halo_r1.md), averaged over the launches of the group at cfg2
    ncu_traffic = {"conv3x3_bwd_weight": (1.0891e9 + 0.8074e9 + 2.4947e9) / 3.0}
    roof = {"bound": "tensor", "kernel": top["kernel"], "achieved": top["tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": top["tflops"] / tf32_peak, "traffic": ncu_traffic.get(top["kernel"]),
            "traffic_note": "bytes per launch from profiles/ (ncu); algorithmic x + dy bytes of the three layers average 1.25e9",
            "peak_source": f"{peaks['source']} bf16_tflops_sustained/2 (TF32 = half the bf16 tensor rate)",
            "share_of_step": top["share"],
            # 3xTF32 issues three tf32 MMAs per algorithmic product, so `frac` tops out at 1/3 for precision-3 kernels;
            # the tensor-pipe occupancy is frac * mma_per_product (ncu: sm__pipe_tensor_cycles_active, profiles/)
            "mma_per_product": 3 if ops.config.conv_wgrad == 3 else 1,
            "tensor_pipe_frac": top["tflops"] / tf32_peak * (3 if ops.config.conv_wgrad == 3 else 1)}
    # BASELINE.json asks for the achieved fraction of the attention-matmul roofline next to the headline number
    att = [g for g in groups if g["kernel"].startswith("sdpa")]
    att_ms, att_gf = sum(g["ms_per_step"] for g in att), sum(g["gflop_per_step"] for g in att)
    mult = 3 if ops.config.attn == 3 else 1
    attention = {"ms_per_step": att_ms, "gflop_per_step": att_gf, "achieved": att_gf / att_ms if att_ms else 0.0, "unit": "TFLOP/s",
                 "peak": tf32_peak, "frac": (att_gf / att_ms / tf32_peak) if att_ms else 0.0, "mma_per_product": mult,
                 "tensor_pipe_frac": (att_gf / att_ms / tf32_peak * mult) if att_ms else 0.0,
                 "note": "QK^T, PV and the four backward products as batched tcgen05 GEMMs incl. the fp32 softmax kernels between "
                         "them; 1.7% of the step's FLOPs in ~60-tile problems, bound by launches and epilogues, not by the pipe"}
    out = {"metric": METRIC, "value": world * B / (ms_dev / 1e3), "unit": "utt/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32" if not args.precision else "f32(" + args.precision + ")", "data": "synthetic",
           "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": world * B, "t_src": T, "t_tgt": cfg.tgt_max_len,
                      "dropout": cfg.dropout, "label_smoothing": cfg.label_smoothing,
                      "precision": {"linear": ops.config.linear, "conv": ops.config.conv, "conv_wgrad": ops.config.conv_wgrad,
                                    "attn": ops.config.attn, "attn_bwd": ops.config.attn_bwd,
                                    "legend": "0 = fp32 CUDA cores, 1 = tcgen05 TF32 (RN), 3 = tcgen05 3xTF32 (fp32-grade)"},
                      "step": "zero_grad+fwd+CE+bwd+allreduce+adam", "parallelism": f"dp{world}",
                      "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush"},
           "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "utt/s", "h2d_bytes_per_step": world * (src_h.numel() * 4 + tgt_h.numel() * 8),
                   "d2h_bytes_per_step": world * 4, "ms_per_step": ms_e2e, "wall_ms_per_step": wall_ms},
           "gpu_launches": launches, "clocks": clocks, "roofline": roof, "attention_roofline": attention, "kernels": groups[:24], "profiled_ms_per_step": ms_prof,
           "final_loss": final_loss}
    if world == 1 and not args.no_cpu:
        ups, mean, threads = cpu_oracle_throughput(2, 2, 1)
        out["cpu_baseline"] = {"value": ups, "unit": "utt/s", "cores": threads, "kind": "port",
                               "sample": "oracle fwd+bwd, 2 utterances of the cfg2 shape per step, 2 timed steps"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (debug only; the metric uses 32)")
    ap.add_argument("--precision", default="", help="linear,conv,attn[,conv_wgrad[,attn_bwd]] in {fp32,tf32,tf32x3} (default: package default)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
