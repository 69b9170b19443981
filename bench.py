"""Benchmark of the DA-CLIP universal-restoration hot path (BASELINE.json metric:
restored images/sec @256^2 T=100; ms per denoiser step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--size S] [--mode posterior|sde]

A "step" is ONE full restoration of one batch: T=100 denoiser evaluations (one CUDA-graph replay each) with the
fused posterior (or SDE) update after each.  Workload = BASELINE.json configs[1]: 256x256, batch 16 per GPU, bf16
tensor-core math with fp32 state, synthetic images and seeded random weights in the reference's state-dict
layout (no checkpoint/dataset offline).  N>1: one process per GPU (torchrun), the batch is sharded by image
(weak scaling: 16 images per GPU), no collective inside the loop, ONE all_gather of the restored shard per step.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import re
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints ONE JSON line on stdout: NCCL's banner / debug output (printed at every level >= VERSION) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

T_STEPS = 100
# algorithmic FLOPs of ONE denoiser evaluation of ONE image (2*MAC, reference shapes).
ALGO_GFLOP_PER_EVAL = {256: 266.2, 512: 1129.1}
# ... of which the layers that run in the tcgen05 implicit-GEMM kernel (all nn.Conv2d / nn.Linear = 257.98 @256^2
# plus the LinearAttention "apply" einsum that is folded into to_out = 1.41): everything except the k-softmax
# context einsum (1.41) and the self-attention QK^T/PV (5.37).  At 512^2: 1031.9 + 5.64.
CONV_GFLOP_PER_EVAL = {256: 257.98 + 1.41, 512: 1031.9 + 5.64}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the tcgen05 GEMM kernels come from the committed summary of
# the ncu launch list of this same command in steady state (tools/ncu_launches.py -> profiles/*_ncu_launch_summary.json:
# `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --launch-skip N -c M python bench.py`);
# the ALGORITHMIC bytes beside them are computed here from the plans (every operand tensor once).


def measured_traffic(B, S):
    """(mean DRAM bytes per launch of the tcgen05 GEMM kernels, source) from the newest committed ncu launch summary
    whose configuration matches, else (None, why)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_launch_summary.json")), reverse=True):
        try:
            d = json.load(open(path))
        except Exception:
            continue
        if d.get("batch") == B and d.get("image") == S and d.get("gemm_kernels"):
            g = d["gemm_kernels"]
            return g["dram_bytes_per_launch"], f"{os.path.relpath(path, ROOT)} ({g['launches']} launches)"
    return None, "no committed ncu launch summary for this configuration"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1386.5), d.get("hbm_gbs", 6538.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def build_product(dev, B, S):
    from daclip_b200 import synthetic
    from daclip_b200.daclip import DaCLIP
    from daclip_b200.model import DenoisingModel
    from daclip_b200.sde import IRSDE
    sd, kw = synthetic.unet_state_dict(0)
    opt = {"gpu_ids": [0], "is_train": False, "dist": False, "model": "denoising",
           "network_G": {"which_model_G": "ConditionalUNet", "setting": dict(kw)},
           "path": {"pretrain_model_G": None, "strict_load": True}}
    model = DenoisingModel(opt)
    model.load_state_dict_into_model(sd)
    sde = IRSDE(max_sigma=50, T=T_STEPS, schedule="cosine", eps=0.005, device=dev)      # options/test.yml:7-12
    sde.set_model(model.model)
    clip = DaCLIP().load_reference_state_dict(synthetic.daclip_visual_state_dict(10)).to(dev).eval()
    return model, sde, clip


def conv_time_per_eval(eng, reps=3):
    """Live CUDA-event timing (on the launching stream) of every conv_igemm launch of one denoiser evaluation,
    run eagerly; returns (sum of conv kernel ms, sum of all kernel ms, number of conv launches)."""
    best = None
    for _ in range(reps):
        evs = []
        for name, fn in eng.steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((name, a, b))
        torch.cuda.synchronize()
        conv = sum(a.elapsed_time(b) for n, a, b in evs if eng.is_conv(n))
        total = sum(a.elapsed_time(b) for n, a, b in evs)
        if best is None or total < best[1]:
            best = (conv, total, sum(1 for n, _, _ in evs if eng.is_conv(n)),
                    sorted(((a.elapsed_time(b), n) for n, a, b in evs), reverse=True)[:12],
                    [(n, a.elapsed_time(b)) for n, a, b in evs])
    return best


# the "ConditionalUNet convs" sub-roofline of the north_star = every nn.Conv2d of the module
# (3x3 205.58 + 1x1 23.89 + 4x4 4.29 + 7x7 2.47 GFLOP per 256^2 image); the spatial subset leaves out the 1x1
# to_qkv / to_out / proj_in / proj_out layers, whose cost is their softmax / LayerNorm / context epilogues.
CONV2D_GFLOP = {256: 236.23}
SPATIAL_GFLOP = {256: 205.58 + 8.05 + 4.29 + 2.47}
_SPATIAL = re.compile(r"block[12]$|res_conv$|^init_conv$|^final_conv$|^(downs|ups)\.\d\.3$")
_POINTWISE = re.compile(r"(?<!attn1)\.(to_q|to_kv|to_qkv|to_q_out|to_out|proj_in|proj_out)$")


def conv_subroofline(layers, B, S, peak_tf):
    sp = sum(ms for n, ms in layers if _SPATIAL.search(n))
    pw = sum(ms for n, ms in layers if _POINTWISE.search(n))
    scale = (S / 256) ** 2 * B
    out = {}
    if sp > 0:
        tf = SPATIAL_GFLOP[256] * scale / sp
        out["spatial_convs"] = {"gflop": round(SPATIAL_GFLOP[256] * scale, 1), "ms": round(sp, 3),
                                "tflops": round(tf, 1), "frac": round(tf / peak_tf, 4)}
    if sp + pw > 0:
        tf = CONV2D_GFLOP[256] * scale / (sp + pw)
        out["all_conv2d"] = {"gflop": round(CONV2D_GFLOP[256] * scale, 1), "ms": round(sp + pw, 3),
                             "tflops": round(tf, 1), "frac": round(tf / peak_tf, 4)}
    return out


def library_baseline(dev, B, S, mode, evals=5):
    """The 'library kernel' bar: the same algorithm as plain PyTorch ops (cuDNN / cuBLAS) ON THE GPU -
    the oracle restatement moved to the device, fp32 and bf16 autocast - timed with CUDA events.  A baseline leg only:
    nothing here is reachable from the product path."""
    from daclip_b200 import synthetic
    from oracle import sde_oracle as So
    from oracle import unet_oracle as O
    sd, kw = synthetic.unet_state_dict(0)
    sd = {k: v.to(dev) for k, v in sd.items()}
    cfg = O.UNetConfig(**kw)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in
           synthetic.restoration_inputs(B, S, S, T=2, seed=100).items()}
    sch = So.Schedule(50, T_STEPS, "cosine", 0.005)
    den = O.make_denoiser(sd, cfg)
    step = So.posterior_step if mode == "posterior" else So.sde_step
    out = {}
    # third leg: the library's best shot - bf16 autocast, channels_last activations and conv weights (NHWC cuDNN kernels),
    # fused scaled_dot_product_attention for the SpatialTransformers
    sd_cl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
    den_cl = O.make_denoiser(sd_cl, cfg)
    for label, ctx, fast in (("fp32", torch.autocast("cuda", enabled=False), False),
                             ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16), False),
                             ("bf16_autocast_channels_last_sdpa", torch.autocast("cuda", dtype=torch.bfloat16), True)):
        try:
            O.USE_SDPA = fast
            net = den_cl if fast else den
            lq = inp["lq"].contiguous(memory_format=torch.channels_last) if fast else inp["lq"]
            x = inp["lq"] + inp["eps0"] * sch.max_sigma
            times = []
            with torch.no_grad(), ctx:
                for i in range(evals + 2):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    xin = x.contiguous(memory_format=torch.channels_last) if fast else x
                    n = net(xin, lq, float(T_STEPS - i), text_context=inp["text_context"],
                            image_context=inp["image_context"])
                    x = step(sch, x, inp["lq"], n.float().contiguous(), inp["noise"][i % 2], T_STEPS - i)
                    b.record()
                    torch.cuda.synchronize()
                    times.append(a.elapsed_time(b))
            ms = sorted(times[2:])[len(times[2:]) // 2]
            out[label] = {"ms_per_denoiser_step": round(ms, 3), "images_per_s": round(B / (ms * T_STEPS / 1e3), 3)}
        except Exception as e:      # a baseline leg must never take the product line down with it
            out[label] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
        finally:
            O.USE_SDPA = False
    out["what"] = (f"oracle restatement as eager PyTorch {torch.__version__} ops (cuDNN/cuBLAS) on the same GPU, "
                   f"batch {B} at {S}x{S}, median of {evals} denoiser+update steps after 2 warm-ups, x{T_STEPS}")
    return out


def run_product(args):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, K, W = args.batch, args.size, args.steps, args.warmup
    from daclip_b200 import lib, synthetic
    lib.load()
    model, sde, clip = build_product(dev, B, S)
    inp = synthetic.restoration_inputs(B, S, S, T=1, seed=100 + rank)
    g = torch.Generator().manual_seed(200 + rank)
    clip_img_host = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    lq_host = inp["lq"].pin_memory()
    lq = lq_host.to(dev)
    img_ctx, deg_ctx = clip.encode_image(clip_img_host.to(dev), control=True)
    from daclip_b200.parallel import gather_restored
    net = model.model.module
    eng = net.engine(B, S, S)
    torch.manual_seed(1234 + rank)

    def restore_resident():
        """Inputs already in HBM: noisy start state -> T-step loop -> (N>1) all_gather of the restored shard."""
        x_T = sde.noise_state(lq)
        model.feed_data(x_T, lq, None, text_context=deg_ctx, image_context=img_ctx)
        model.test(sde, mode=args.mode)
        return gather_restored(model.output, world * B)     # one NCCL all_gather of the restored shard (N > 1)

    out_host = torch.empty(B, 3, S, S).pin_memory()

    def restore_e2e():
        """The reference's per-batch user flow (test.py:112-127) with HOST buffers: H2D of the LQ batch and the
        224^2 CLIP views, encode_image(control=True), noise_state, feed_data, test, D2H of the restored batch."""
        lq_d = lq_host.to(dev, non_blocking=True)
        ci_d = clip_img_host.to(dev, non_blocking=True)
        ic, dc = clip.encode_image(ci_d, control=True)
        x_T = sde.noise_state(lq_d)
        model.feed_data(x_T, lq_d, None, text_context=dc, image_context=ic)
        model.test(sde, mode=args.mode)
        out_host.copy_(model.output, non_blocking=True)
        torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        restore_resident()
    sync_all()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        start.record()
        for _ in range(K):
            restore_resident()
        end.record()
        sync_all()
    ms = start.elapsed_time(end)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    ms_per_step = ms / K
    value = world * B * K / (ms / 1e3)

    # end-to-end through the public API with host buffers
    for _ in range(max(1, min(W, 2))):
        restore_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        restore_e2e()
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / e2e_s.item()

    result = None
    if rank == 0:
        conv_ms, all_ms, n_conv, top, layers = conv_time_per_eval(eng)
        if args.dump_layers:
            with open(args.dump_layers, "w") as f:
                for n, ms_ in layers:
                    fl = getattr(eng, "layer_flops", {}).get(n)
                    tf = f"{fl / 1e9:8.1f} GFLOP {fl / ms_ / 1e9:7.1f} TFLOP/s" if fl else ""
                    f.write(f"{ms_ * 1e3:9.1f} us  {'conv' if eng.is_conv(n) else 'aux '}  {n:28s} {tf}\n")
        peak_tf, peak_bw, which = peaks()
        gflop = CONV_GFLOP_PER_EVAL.get(S, CONV_GFLOP_PER_EVAL[256] * (S / 256) ** 2) * B
        achieved = gflop / n_conv / (conv_ms / n_conv) if conv_ms > 0 else 0.0       # GFLOP/ms == TFLOP/s
        launches_per_eval = eng.launches + 2                                          # + loop tick + fused SDE update
        n_lin = len(getattr(eng, "linattn_names", ()))
        lb = getattr(eng, "layer_bytes", {})
        algo_bytes = sum(lb.values()) / max(1, len(lb))                               # mean per tcgen05 GEMM launch
        traffic, traffic_src = measured_traffic(B, S)
        algo = ALGO_GFLOP_PER_EVAL.get(S, ALGO_GFLOP_PER_EVAL[256] * (S / 256) ** 2)
        result = {
            "metric": "restored images/sec @256^2 T=100", "value": round(value, 3), "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_per_step, 3),
            "ms_per_denoiser_step": round(ms_per_step / T_STEPS, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{S}x{S} batch {B} per GPU, T={T_STEPS} {args.mode} sampling, IR-SDE + "
                                   f"ConditionalUNet(nf64, ch_mult 1-2-4-8) + DA-CLIP contexts; random weights in "
                                   f"the reference state-dict layout",
                       "global_batch": world * B, "image": S, "T": T_STEPS, "sampling_mode": args.mode,
                       "parallelism": f"batch-sharded x{world}, one all_gather per restoration",
                       "l2": "working set per denoiser step (>2 GB of activations) exceeds the 126 MB L2"},
            "whole_step_tensor_frac": round(world and (algo * B * T_STEPS / ms_per_step) / peak_tf, 4),
            "roofline": {"bound": "tensor",
                         "kernel": "the tcgen05 GEMM launches of one evaluation: conv_igemm_kernel (every Conv2d / Linear) "
                                   "and the chained LinearAttention kernels linattn_kv(2)_kernel / linattn_qout(2)_kernel",
                         "kernels_summed": {"conv_igemm_kernel": n_conv - n_lin, "linattn_kv/qout kernels": n_lin},
                         "achieved": round(achieved, 2), "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": round(achieved / peak_tf, 4),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": round(algo_bytes, 1),
                         "peak_source": f"{which} bf16_tflops_sustained",
                         "launches_per_eval": n_conv, "conv_ms_per_eval": round(conv_ms, 3),
                         "all_kernels_ms_per_eval": round(all_ms, 3),
                         "algorithmic_gflop_per_eval": round(gflop, 1),
                         "slowest": [[round(a, 3), n] for a, n in top]},
            "e2e": {"value": round(e2e_value, 3), "unit": "images/s",
                    "h2d_bytes_per_step": int(lq_host.numel() * 4 + clip_img_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4),
                    "includes": "H2D, DaCLIP.encode_image(control=True), noise_state, T-step loop, D2H"},
            "gpu_launches": int(K * T_STEPS * launches_per_eval),
            "clocks": clocks.summary(),
        }
        result["roofline"]["unet_convs"] = conv_subroofline(layers, B, S, peak_tf)
        if world == 1 and not args.no_library_baseline:
            result["library_baseline"] = library_baseline(dev, B, S, args.mode)
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(S, args.mode, evals=args.cpu_evals)
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_sample(S, mode, evals, threads=None):
    """ONE protocol for both CPU legs (`cpu_baseline` of the product line and `--impl reference`): batch 1 (the reference's
    test loader is hard-wired to batch 1, data/__init__.py:30-33), `evals` denoiser + update steps through the sampler's
    own loop after a one-step warm-up, all host threads, fp32.  Runs THE REFERENCE's modules (IRSDE.reverse_* driving
    ConditionalUNet, vendored unmodified into the git-ignored baseline/_ref/ by oracle/vendor_reference.py) when they
    are present, else the oracle port.  Returns (seconds per denoiser step, threads, kind)."""
    from daclip_b200 import synthetic
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd, kw = synthetic.unet_state_dict(0)
    inp = synthetic.restoration_inputs(1, S, S, T=1, seed=100)
    from oracle import vendor_reference as V
    if V.available():
        IRSDE, UNet = V.load_reference()
        net = UNet(**kw)
        net.load_state_dict(sd, strict=True)
        net.eval()
        sde = IRSDE(max_sigma=50, T=T_STEPS, schedule="cosine", eps=0.005, device="cpu")     # options/test.yml:7-12
        sde.set_model(net)
        sde.set_mu(inp["lq"])
        x = sde.noise_state(inp["lq"])
        loop = sde.reverse_posterior if mode == "posterior" else sde.reverse_sde
        run = lambda k: loop(x, T=k, text_context=inp["text_context"], image_context=inp["image_context"])
        kind = "reference"
    else:
        from oracle import sde_oracle as So
        from oracle import unet_oracle as O
        sch = So.Schedule(50, T_STEPS, "cosine", 0.005)
        den = O.make_denoiser(sd, O.UNetConfig(**kw))
        x = inp["lq"] + inp["eps0"] * sch.max_sigma
        run = lambda k: So.reverse(sch, den, x, inp["lq"], mode=mode, T=k, text_context=inp["text_context"],
                                   image_context=inp["image_context"])
        kind = "port"
    with torch.no_grad():
        run(1)
        t0 = time.perf_counter()
        run(evals)
        per_eval = (time.perf_counter() - t0) / evals
    return per_eval, threads, kind


def cpu_baseline(S, mode, evals):
    per_eval, threads, kind = cpu_sample(S, mode, evals)
    what = ("the reference's IRSDE + ConditionalUNet (baseline/_ref, unmodified)" if kind == "reference"
            else "oracle port (fp32 PyTorch restatement of the reference)")
    return {"value": round(1.0 / (per_eval * T_STEPS), 5), "unit": "images/s", "cores": threads, "kind": kind,
            "s_per_denoiser_step": round(per_eval, 3),
            "sample": f"{what}, batch 1 at {S}x{S}: {evals} of T={T_STEPS} denoiser+{mode} steps through the sampler loop "
                      f"after a 1-step warm-up, extrapolated x{T_STEPS}/{evals}"}


def run_reference(args):
    """Reference arm: the reference's own implementation of the path on the host CPU (baseline/_ref, vendored unmodified
    from /root/reference by oracle/vendor_reference.py; the oracle port if that directory is missing).  Same protocol
    as `cpu_baseline`; each bench step is a bounded sample (--cpu-evals denoiser steps of one image), extrapolated to
    images/s at T=100."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, S, K, W = args.batch, args.size, args.steps, args.warmup
    evals = max(1, min(args.cpu_evals, 3))
    vals, kind, threads = [], "port", os.cpu_count()
    for _ in range(W + K):
        per_eval, threads, kind = cpu_sample(S, args.mode, evals)
        vals.append(per_eval)
    per_eval = sum(vals[W:]) / K
    value = 1.0 / (per_eval * T_STEPS)
    sample = (f"batch 1, {evals} denoiser+{args.mode} steps per bench step through the "
              f"{'reference' if kind == 'reference' else 'oracle'} sampler loop (+1 warm-up step each), x{T_STEPS}/{evals}")
    out = {"impl": "reference", "metric": "restored images/sec @256^2 T=100", "value": round(value, 5),
           "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K, "warmup": W,
           "ms_per_step": round(per_eval * T_STEPS * 1e3 * B, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{S}x{S} batch {B}, T={T_STEPS} {args.mode} sampling (CPU sample: {sample})",
                      "image": S, "T": T_STEPS},
           "cpu_baseline": {"value": round(value, 5), "unit": "images/s", "cores": threads, "kind": kind,
                            "sample": sample},
           "e2e": {"value": round(value, 5), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--mode", default="posterior", choices=["posterior", "sde"])
    ap.add_argument("--cpu-evals", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip timing the oracle restatement as eager PyTorch (cuDNN/cuBLAS) on the GPU")
    ap.add_argument("--library-baseline", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--dump-layers", default=None, help="write the per-launch CUDA-event times of one evaluation here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    run_product(args)


if __name__ == "__main__":
    main()
