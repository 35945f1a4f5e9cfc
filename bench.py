#!/usr/bin/env python
"""Benchmark of the CFG-DDPM sampling hot path (BASELINE.json metric: spectrograms/sec, full T-step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2] -- full CFG sampling (cond+uncond batched 2x) of n=512
spectrograms per GPU at the reference's generation size (latent [n,4,64,64], 27 classes, cfg_scale=3,
noise_steps=1000), 16-bit tensor-core engine, CUDA-graph loop.  One bench "step" = ONE denoising timestep of
that loop over the whole batch: 2n UNet rows + CFG lerp + posterior update + Philox noise.  A spectrogram
costs noise_steps-1 = 999 such steps, so  spectrograms/s = n_total / (999 * seconds_per_step).
The activations of one step (several GB) are far larger than the 126 MB L2, so no explicit L2 flush is needed.

One JSON line is printed by rank 0.  Extra keys: roofline (dominant kernel family, timed live with CUDA
events in an eager per-launch pass), kernels (per-family time shares), cpu_baseline (oracle port on the
host cores, bounded sample), e2e (through Diffusion.sample with pinned host labels in / uint8 host images out).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_STEPS = 1000
NUM_CLASSES = 27
METRIC = "cfg_ddpm_spectrograms_per_sec"
UNIT = "spectrograms/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"],
                "source": "measured (MEASURED_PEAKS.json; tensor = sustained bf16)"}
    return {"hbm": 6650.0, "tensor": 1400.0, "tensor_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,enforced.power.limit")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw, plim = [], None, set(), [], None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
            try:  # board power during the timed region next to the enforced limit: the sustained loop is power-bound
                pw.append(float(f[2]))
                plim = float(f[7]) if len(f) > 7 else plim
            except ValueError:
                pass
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons), "power_w": pw[len(pw) // 2] if pw else None, "power_limit_w": plim}


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores (bounded sample)
# ----------------------------------------------------------------------------------------------------
def cpu_baseline(S, c, n_cpu, steps, warmup=1):
    from oracle import ddpm_oracle as O
    from oracle.weights import make_state_dict

    torch.set_num_threads(os.cpu_count())
    sd = make_state_dict(1234, c, c, NUM_CLASSES)
    y = torch.arange(n_cpu) % NUM_CLASSES
    g = torch.Generator().manual_seed(0)
    noise = torch.randn((warmup + steps + 1, n_cpu, c, S, S), generator=g)
    ts = []
    beta, alpha, ah = O.noise_schedule(T_STEPS)
    c1, c2, c3 = O.posterior_coefficients(beta, alpha, ah)
    x = noise[0]
    for k in range(warmup + steps):
        i = T_STEPS - 1 - k
        t0 = time.perf_counter()
        t = (torch.ones(n_cpu) * i).long()
        e_c = O.unet_forward(sd, x, t, y)
        e_u = O.unet_forward(sd, x, t, None)
        eps = O.cfg_combine(e_c, e_u, 3)
        x = O.posterior_update(x, eps, c1[i], c2[i], c3[i], noise[k + 1])
        ts.append(time.perf_counter() - t0)
    sec = sum(ts[warmup:]) / steps
    return {
        "value": n_cpu / (sec * (T_STEPS - 1)),
        "unit": UNIT,
        "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": f"{steps} CFG timesteps (2 UNet fwd + update each) of n={n_cpu} at [{c},{S},{S}] fp32 on the host, "
                  f"{sec:.3f} s/step, extrapolated x{T_STEPS - 1} timesteps; oracle port of the reference "
                  "(same ATen CPU kernels, attention without the discarded head-averaged weights)",
        "sec_per_step": sec,
    }


# ----------------------------------------------------------------------------------------------------
# algorithmic FLOPs / bytes per launch family
# ----------------------------------------------------------------------------------------------------
def classify(fn, a, kw):
    """(family, algorithmic flops, algorithmic bytes) of one recorded plan op."""
    from spectrogramgenai_b200 import ops

    name = fn.__name__
    if name == "igemm_launch":
        g = a[0]
        M = g.rows * g.H * g.W
        fl = 2.0 * M * g.Cin * g.Cout * g.taps
        esz = 4 if g.act_dtype == 0 else 2
        by = M * g.Cin * esz + g.taps * g.Cin * g.Cout * esz + M * g.Cout * (4 if g.out_f32 else 0) \
            + M * g.Cout * ((2 if g.out_dtype else esz) if g.out_act else 0) + (M * g.Cout * 4 if g.residual else 0)
        eng = "tc" if g.engine == 1 else "simt"
        return (f"igemm_{eng}_conv3x3" if g.taps == 9 else f"igemm_{eng}_linear"), fl, by
    if name == "attention":
        rows, L, C = kw["rows"], kw["L"], kw["C"]
        return "attention", 4.0 * rows * L * L * C, rows * L * C * (3 * a[0].element_size() + a[1].element_size())
    if name == "gn_apply":
        raw = a[0]
        by = raw.numel() * raw.element_size()
        for k in ("out_f32", "out_act", "residual"):
            t = kw.get(k)
            if t is not None:
                by += t.numel() * t.element_size()
        return "gn_apply", 0.0, by
    if name == "gn_apply_vcat":
        raw, part, gamma, beta, xs, skip, out = a
        return "gn_apply", 0.0, raw.numel() * 2 + xs.numel() * 4 + skip.numel() * 4 * (raw.shape[0] // skip.shape[0]) \
            + out.numel() * out.element_size()
    if name in ("maxpool2", "upsample_cat"):
        by = sum(t.numel() * t.element_size() for t in a)
        for k in ("out_f32", "out_act"):
            t = kw.get(k)
            if t is not None:
                by += t.numel() * t.element_size()
        return name, 0.0, by
    if name == "layernorm":
        return "layernorm", 0.0, a[0].numel() * 4 + a[3].numel() * a[3].element_size()
    if name == "ln_inproj":  # x fp32 in, qkv 16-bit out; one [M,C] x [C,3C] GEMM
        x, qkv = a[0], a[5]
        Cc = x.shape[-1]
        return "sa_ln_inproj_fused", 2.0 * x.numel() * 3 * Cc, x.numel() * 4 + qkv.numel() * qkv.element_size()
    if name == "attn_tail":  # att 16-bit + x fp32 in, out fp32; three [M,C] x [C,C] GEMMs
        att, x, out = a[0], a[1], a[10]
        Cc = x.shape[-1]
        fl, by = 3 * 2.0 * x.numel() * Cc, att.numel() * att.element_size() + x.numel() * 4
        if out is not None:
            by += out.numel() * 4
        if kw.get("outc") is not None:  # fused 1x1 output conv: eps NCHW written instead of (or besides) the block output
            eps = kw["outc"][2]
            fl += 2.0 * eps.numel() * Cc
            by += eps.numel() * 4
        return "sa_tail_fused", fl, by
    if name == "conv_in":
        x, w, raw, part = a
        return "conv_in", 2.0 * raw.numel() * x.shape[1] * 9, raw.shape[0] * x[0].numel() * 4 + raw.numel() * 4
    if name == "conv_out":
        x, w, b, eps = a
        return "conv_out", 2.0 * x.numel() * eps.shape[1], x.numel() * 4 + eps.numel() * 4
    if name == "time_embed":
        return "time_embed", 0.0, a[-1].numel() * 4 + a[-2].numel() * 4
    return name, 0.0, 0.0


def per_launch_profile(plan, extra_ops, reps=2):
    """Eager pass with a CUDA event pair around every launch on the launching (current) stream."""
    ops_list = list(plan.ops) + extra_ops
    acc = {}
    for rep in range(reps + 1):
        evs = []
        torch.cuda.synchronize()
        for fn, a, kw in ops_list:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **kw)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if rep == 0:
            continue  # warm-up
        for (fn, a, kw), (e0, e1) in zip(ops_list, evs):
            fam, fl, by = classify(fn, a, kw)
            d = acc.setdefault(fam, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += e0.elapsed_time(e1) / reps
            d["flops"] += fl / reps
            d["bytes"] += by / reps
            d["launches"] += 1.0 / reps
    return acc


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.size, args.channels, args.cpu_batch, max(1, args.steps), max(1, min(args.warmup, 2)))
    out = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.cpu_batch, "fp32 CPU"),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def workload_config(args, n, engine):
    if (args.size, args.channels) == (64, 4):
        which = "BASELINE configs[2]"  # the configuration the metric is quoted on
    elif (args.size, args.channels) == (256, 1):
        which = "BASELINE configs[4] (pixel-space 256x256, attention-dominated)"
    else:
        which = "non-baseline geometry"
    return {
        "workload": f"{which}: full CFG-DDPM sampling loop, cond+uncond batched 2x, cfg_scale=3, "
                    f"noise_steps={T_STEPS}, UNet input [{args.channels},{args.size},{args.size}], {NUM_CLASSES} classes",
        "batch_per_gpu": n, "engine": engine,
        "step": "one denoising timestep over the batch (2n UNet rows + CFG lerp + posterior update + noise)",
        "timesteps_per_spectrogram": T_STEPS - 1,
        "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush",
        "loop": "CUDA graph replay per timestep",
    }


_JSON_OUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner to stdout
    when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the whole process and
    the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="spectrograms per GPU")
    ap.add_argument("--size", type=int, default=64, help="UNet input size (reference generation config: 256/4 = 64)")
    ap.add_argument("--channels", type=int, default=4)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "f16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-steps", type=int, default=10, help="CPU-baseline sample: timesteps of n = cpu-batch (~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    # stdout carries exactly ONE JSON line: NCCL's own banner / debug output (it writes to stdout) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

    from spectrogramgenai_b200 import ops
    from spectrogramgenai_b200.diff_modules import Diffusion
    from spectrogramgenai_b200.sharding import gather_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, S, c, K, Wm = args.batch, args.size, args.channels, args.steps, args.warmup
    peaks = load_peaks()

    torch.manual_seed(42)  # the reference's seed (ddpm_conditional_generate.py:22): synthetic random-init weights
    d = Diffusion(noise_steps=T_STEPS, img_size=S, num_classes=NUM_CLASSES, c_in=c, c_out=c, device=dev,
                  compute_dtype=args.mode)
    model = d.model
    plan = model.plan(n_src=n, rows=2 * n, S=S, use_step=True)
    x = plan.x_in
    labels = (torch.arange(n) + rank * n) % NUM_CLASSES
    plan.y.fill_(-1)
    plan.y[:n].copy_(labels.to(dev))
    seed, base = 42, rank * n

    def one_step():
        plan.run()
        ops.cfg_update(x, plan.eps, d._coef, plan.step, cfg_scale=3.0, seed=seed, sample_base=base)
        ops.step_advance(plan.step)

    launches_per_step = plan.n_launches + 2
    ops.philox_normal(x, seed=seed, sample_base=base, step_tag=T_STEPS)
    plan.step.fill_(T_STEPS - 1)
    one_step()  # eager: module loading / attribute setting outside capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one_step()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed K steps (inputs resident in HBM) ----
    for _ in range(Wm):
        graph.replay()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / K
    value = world * n / (ms_per_step * 1e-3 * (T_STEPS - 1))
    finite = bool(torch.isfinite(x).all().item())

    # ---- e2e: public API, pinned host labels in, uint8 host images out, the one NCCL gather included ----
    labels_host = labels.clone().pin_memory()
    out_host = torch.empty((world * n if rank == 0 or world > 1 else n, c, S, S), dtype=torch.uint8).pin_memory()
    d.sample(False, labels_host, 3, seed=seed, sample_base=base, micro_batch=n, max_steps=1)  # warm the API path
    sync_all()
    t0 = time.perf_counter()
    u8 = d.sample(False, labels_host, 3, seed=seed, sample_base=base, micro_batch=n, max_steps=K)
    if world > 1:
        u8 = gather_shards(u8, world * n)
    out_host[: u8.shape[0]].copy_(u8, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    e2e_value = world * n / ((e2e_s / K) * (T_STEPS - 1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-launch profile (eager, CUDA events on the launching stream) -> roofline of the dominant family ----
    kernels, roofline = {}, None
    if not args.no_profile:
        extra = [(ops.cfg_update, (x, plan.eps, d._coef, plan.step), dict(cfg_scale=3.0, seed=seed, sample_base=base))]
        plan.step.fill_(T_STEPS - 1)
        acc = per_launch_profile(plan, extra)
        tot = sum(v["ms"] for v in acc.values())
        for fam, v in sorted(acc.items(), key=lambda kv: -kv[1]["ms"]):
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0
            gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
            kernels[fam] = {"ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 4), "launches": round(v["launches"]),
                            "tflops": round(tf, 2), "gbs": round(gb, 1)}
        kernels["_eager_step_ms"] = round(tot, 3)
        dom = max((f for f in acc if not f.startswith("_")), key=lambda f: acc[f]["ms"])
        v = acc[dom]
        tensor_bound = v["flops"] / max(v["bytes"], 1.0) > peaks["tensor"] * 1e12 / (peaks["hbm"] * 1e9)
        if tensor_bound:
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tensor"],
                        "unit": "TFLOP/s", "frac": round(ach / peaks["tensor"], 4), "traffic": None}
        else:
            ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm"],
                        "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4), "traffic": None}
        # DRAM traffic per launch of the dominant family from the committed `ncu --set full` capture of this command
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic_r1.json")) as f:
                tr = json.load(f).get(dom)
            if tr and int(tr.get("batch_per_gpu", -1)) == n and tr.get("mode") == args.mode:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
        except (OSError, ValueError):
            pass
        if dom == "attention":
            # d = 16 heads: a score tile is 0.5 M tensor MACs but 16 K exponentials, so the binding unit is the MUFU
            # pipe (16 ex2/clk/SM, measured 15.8), not the tensor pipe: report that roofline beside the tensor one
            exps = sum(float(kw["rows"]) * 4 * kw["L"] * kw["L"] for fn, a, kw in plan.ops if fn.__name__ == "attention")
            sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
            mufu_peak = 16.0 * 148 * sm_mhz * 1e6
            roofline["binding_unit"] = {
                "unit": "MUFU ex2 (16/clk/SM at the sampled SM clock); 1/4 of the exponentials run on the FMA pipe",
                "exp_per_s": round(exps / (v["ms"] * 1e-3), 1), "mufu_peak_exp_per_s": mufu_peak,
                "frac_all_on_mufu": round(exps / (v["ms"] * 1e-3) / mufu_peak, 4)}
        roofline["peak_source"] = peaks["source"]
        roofline["launches_per_step"] = round(v["launches"])
        roofline["avg_launch_ms"] = round(v["ms"] / max(v["launches"], 1), 4)

    cb = None
    if not args.no_cpu_baseline:
        cb = cpu_baseline(S, c, args.cpu_batch, args.cpu_steps)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    from oracle.ddpm_oracle import flops_per_forward

    flops_step = 2.0 * n * flops_per_forward(S, c, c)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.mode, "data": "synthetic (random-init weights seed 42, Philox x_T and noise)",
        "config": workload_config(args, n, f"{args.mode} operands, fp32 accumulate (tcgen05)" if args.mode != "fp32" else "fp32 SIMT"),
        "unet_fwd_ms_per_step": ms_per_step,
        "model_tflops": round(world * flops_step / (ms_per_step * 1e-3) / 1e12, 2),
        "model_tflops_frac_of_peak": round(flops_step / (ms_per_step * 1e-3) / 1e12 / peaks["tensor"], 4),
        "finite": finite,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": labels_host.numel() * 8 / K,
                "d2h_bytes_per_step": n * c * S * S / K,
                "note": f"Diffusion.sample(labels on pinned host, max_steps={K}) -> uint8 on pinned host; one call "
                        "includes x_T Philox init, graph capture, K timesteps, uint8 tail"
                        + (", NCCL all_gather of the uint8 output" if world > 1 else "")
                        + "; per-call copies are amortised over K (a real run amortises them over 999)"},
        "gpu_launches": world * K * launches_per_step,
        "launches_per_step": launches_per_step,
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cb,
        "activation_bytes_allocated": plan.nbytes,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
