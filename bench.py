#!/usr/bin/env python
"""bench.py -- deep-image-prior hot path on B200: optimisation iterations/sec, 512x512 skip-net denoising.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one optimisation iteration of BASELINE.json config[1] (denoising F16-sized 512x512, skip[128x5], fp32):
  z = z0 + N(0,1)/30  ->  out = net(z)  ->  MSE(out, noisy target)  ->  backward  ->  Adam(lr 0.01) step.
One independent image per GPU (weak scaling, no data-path collective; NCCL only gathers the result records).

`value`  : iterations/sec summed over ranks with inputs resident in HBM (closure-free device runner, dip_run_iterations)
`e2e`    : the same metric through the notebook-facing API (models.get_net + utils.optimize-style closure loop) with the
           step's perturbed input copied host(pinned)->device and the loss read back device->host inside the timed region
`roofline`: dominant kernel (tcgen05 implicit-GEMM conv, tc_conv_kernel) -- algorithmic FLOPs / CUDA-event device time
`cpu_baseline` / --impl reference: the oracle port of the reference's torch-CPU path on the host cores.
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
sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))

H = W = 512
IN_CH, OUT_CH = 32, 3
SIGMA_REG = 1.0 / 30.0
LR = 0.01
ITERS_PER_IMAGE = 2000           # BASELINE.json config[0]/[1]
ALG_GFLOP_PER_ITER = 460.07      # SURVEY.md section 6 (2*M*N*K over the 26 convs, fwd+dgrad+wgrad)
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch from the committed `ncu --set full` capture
# (profiles/r01_ncu_conv_l0up_v6.txt); None until captured for the current kernel version
ROOFLINE_TRAFFIC_BYTES = 232083456  # 140.32 MB read + 91.76 MB written (profiles/r01_ncu_conv_l0up_v6.txt; algorithmic: 143.2 MB in + 0.8 MB weights + 134.2 MB out)
ROOFLINE_HBM_TRAFFIC_BYTES = 370134272  # 268.58 MB read + 101.55 MB written (profiles/r01_ncu_bn_bwd_apply_l0.txt)
METRIC = "optimisation iterations/sec (512x512 skip-net denoising, sum over independent images)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
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
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_problem(torch, seed):
    """Synthetic inputs of the BASELINE shape (no dataset on the box): z0 ~ U[0,0.1), target = clip(img + N(0,25/255))."""
    g = torch.Generator().manual_seed(1000 + seed)
    z0 = torch.rand(1, IN_CH, H, W, generator=g) * 0.1
    clean = torch.rand(1, OUT_CH, H // 8, W // 8, generator=g)
    clean = torch.nn.functional.interpolate(clean, size=(H, W), mode="bilinear", align_corners=False)
    target = (clean + torch.randn(clean.shape, generator=g) * (25.0 / 255.0)).clamp(0, 1)
    return z0, clean, target


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def best_cpu_threads(torch):
    """torch-CPU gets slower, not faster, when all 100+ logical cores of the GPU box are used (oversubscription of
    MKL-DNN on a shared host): pick the thread count that gives the reference its best iteration time (quarter-size
    probe, one iteration each).  The chosen count is what `cores` reports."""
    from oracle import dip_oracle as O
    cores = os.cpu_count() or 1
    cands = sorted(set(c for c in (8, 16, 32, 64, cores) if c <= cores))
    cfg = O.SkipConfig(upsample_mode="bilinear")
    params = O.init_params(cfg, seed=0)
    z = torch.rand(1, IN_CH, 256, 256) * 0.1
    t = torch.rand(1, OUT_CH, 256, 256)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        for rep in range(2):   # first repetition warms the thread pool
            t0 = time.perf_counter()
            torch.autograd.grad(O.mse_loss(O.skip_forward(params, z, cfg), t), params)
            dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    return best


def cpu_iterations(torch, n_timed, n_warm, threads):
    """Times the oracle port of the reference's per-iteration path on the host cores. Returns (it/s, seconds/iter)."""
    from oracle import dip_oracle as O
    torch.set_num_threads(threads)
    cfg = O.SkipConfig(upsample_mode="bilinear")
    params = O.init_params(cfg, seed=0)
    z0, _, target = make_problem(torch, 0)
    opt = O.Adam(params, LR)
    gen = torch.Generator().manual_seed(123)
    times = []
    for i in range(n_warm + n_timed):
        t0 = time.perf_counter()
        z = z0 + torch.randn(z0.shape, generator=gen) * SIGMA_REG
        out = O.skip_forward(params, z, cfg)
        loss = O.mse_loss(out, target)
        grads = torch.autograd.grad(loss, params)
        opt.step(grads)
        loss.item()
        if i >= n_warm:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return len(times) / total, total / len(times)


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = best_cpu_threads(torch)
    budget_s = float(os.environ.get("DIP_REF_BUDGET_S", "150"))
    # bounded sample: one step = one full-size iteration (~2 s on 8 cores); cap the count so the run ends in minutes
    t0 = time.perf_counter()
    _, s_per = cpu_iterations(torch, 1, 1, cores)
    warm = min(args.warmup, 3)
    steps = max(1, min(args.steps, int((budget_s - (time.perf_counter() - t0)) / s_per) - warm))
    its, s_per = cpu_iterations(torch, steps, warm, cores)
    sample = "%d timed iterations (of %d requested) after %d warm-up, full 512x512 workload, %d threads (best of a probe; host has %d logical cores)" % (
        steps, args.steps, warm, cores, os.cpu_count() or 1)
    line = {"impl": "reference", "metric": METRIC, "value": its, "unit": "it/s", "n_gpus": args.gpus, "steps": steps,
            "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1000.0 * s_per, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "denoise 512x512 skip[128x5] in32 out3 bilinear, noise+fwd+MSE+bwd+Adam per step",
                       "impl_detail": "oracle/dip_oracle.py: the reference's graph on torch-CPU (MKL-DNN), all host cores"},
            "cpu_baseline": {"value": its, "unit": "it/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": its, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dip_engine as de
    import multi_gpu as mg
    import models
    from oracle import dip_oracle as O  # cpu_baseline leg + PSNR helper only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("DIP_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = load_peaks()

    # ---- build the network through the public API (same seeds on every rank; different image per rank)
    torch.manual_seed(0)
    net = models.get_net(IN_CH, "skip", "reflection", skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5,
                         upsample_mode="bilinear").type(torch.cuda.FloatTensor)
    z0_h, clean_h, target_h = make_problem(torch, rank)
    z0, target = z0_h.to(dev), target_h.to(dev)
    params = [p for p in net.parameters()]
    opt = de.FusedAdam(params, lr=LR)

    # one notebook-style step to create the plan, bind parameters/gradients and attach .grad views
    mse = torch.nn.MSELoss()

    def api_step(z_dev):
        opt.zero_grad()
        out = net(z_dev)
        loss = mse(out, target)
        loss.backward()
        opt.step()
        return loss

    api_step(z0)
    torch.cuda.synchronize()
    plan = list(net._dip_plans.values())[0]
    grads = [p.grad for p in params]
    opt._bind(grads)
    out_buf = torch.empty(1, OUT_CH, H, W, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_steps(n, hist=None):
        de.run_iterations(plan, opt, z0, target, None, SIGMA_REG, 1234 + rank, n, LR, out=out_buf, loss_hist=hist)

    # ---- `value`: device-resident runner ---------------------------------------------------------------------
    device_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    hist = torch.zeros(args.steps, dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    device_steps(args.steps, hist)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # roofline pass: the timed region above replays a CUDA graph (no per-kernel events possible inside it), so the
    # tensor-core launches are bracketed with CUDA events in a short eager pass of the same iterations right after it
    roof_steps = min(args.steps, 10)
    os.environ["DIP_NO_SIDE"] = "1"   # kernels timed one at a time (the timed region overlaps the wgrad chain on a side stream)
    plan.set_timing(True)
    device_steps(roof_steps)
    torch.cuda.synchronize()
    records = plan.get_timing_records()
    plan.set_timing(False)
    os.environ.pop("DIP_NO_SIDE", None)
    fwd_l, bwd_l = plan.num_launches()
    launches_per_step = fwd_l + bwd_l + 3      # + noise, mse, adam
    ms_max = mg.max_over_ranks(ms, device=dev)
    value = mg.aggregate_rate(args.steps, ms_max / 1000.0, world)

    # ---- `e2e`: notebook-facing API, host buffers in the timed region -----------------------------------------
    pool = 4
    gen = torch.Generator().manual_seed(77 + rank)
    z_host = [(z0_h + torch.randn(z0_h.shape, generator=gen) * SIGMA_REG).pin_memory() for _ in range(pool)]
    # every step's input crosses PCIe inside the timed region; the copy of step i+1 is issued on a copy stream while
    # step i computes (double-buffered device input), as any input pipeline would do
    z_dev = [torch.empty_like(z0), torch.empty_like(z0)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_steps = args.steps

    def prefetch(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])               # the step that last read this buffer has finished
            z_dev[b].copy_(z_host[i % pool], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_step(i):
        b = i % 2
        torch.cuda.current_stream().wait_event(ready[b])
        loss = api_step(z_dev[b])
        consumed[b].record()
        if i + 1 < n_total:
            prefetch(i + 1)
        return loss.item()                                     # D2H: the step's loss (sync)

    for b in range(2):
        consumed[b].record()
    n_total = 3
    prefetch(0)
    for i in range(3):
        e2e_step(i)
    barrier()
    n_total = e2e_steps
    e0.record()
    prefetch(0)                                                # H2D of step 0 is inside the timed region too
    last = 0.0
    for i in range(e2e_steps):
        last = e2e_step(i)
    e1.record()
    barrier()
    e2e_value = mg.aggregate_rate(e2e_steps, mg.max_over_ranks(e0.elapsed_time(e1), device=dev) / 1000.0, world)

    # ---- result record per rank (the only collective of the job)
    with torch.no_grad():
        out_np = net(z0).cpu().numpy()[0]
    recs = mg.gather_records([O.psnr(clean_h.numpy()[0], out_np), float(hist[-1].item()), args.steps / (ms / 1000.0)],
                             device=dev)

    if rank == 0:
        def agg(pred):
            sel = [r for r in records if pred(r)]
            ms_ = sum(r[2] for r in sel)
            fl_ = sum(r[1] for r in sel)
            return ms_, fl_, len(sel), (fl_ / (ms_ / 1000.0) / 1e12 if ms_ > 0 else 0.0)
        conv_ms, conv_fl, conv_n, conv_all = agg(lambda r: r[0] in (0, 1))
        wg_ms, wg_fl, wg_n, wg_ach = agg(lambda r: r[0] == 2)
        # dominant launch = the largest single tensor-core launch of the step (level-0 3x3 conv 132->128 at 512x512,
        # 79.7 algorithmic GFLOP): its fprop instances
        big = max(r[1] for r in records if r[0] == 0)
        dom_ms, dom_fl, dom_n, achieved = agg(lambda r: r[0] == 0 and r[1] == big)
        tf32_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2.0
        # dominant HBM-bound launch: BatchNorm+LeakyReLU backward (apply pass) of the level-0 3x3 up conv, 128 ch @512x512
        hbm = [r for r in records if r[0] == 3]
        hbm_big = max((r[1] for r in hbm), default=0.0)
        hbm_sel = [r for r in hbm if r[1] == hbm_big]
        hbm_ms = sum(r[2] for r in hbm_sel)
        hbm_gbs = (hbm_big * len(hbm_sel) / (hbm_ms / 1000.0) / 1e9) if hbm_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32 storage; tf32 tensor-core multiplies with fp32 accumulate in the convs (cuDNN's default fp32 mode)",
            "data": "synthetic",
            "config": {"workload": "denoise 512x512 skip[128x5] in32 out3 bilinear, noise+fwd+MSE+bwd+Adam per step "
                                   "(BASELINE.json configs[1]); one independent image per GPU",
                       "iters_per_image": ITERS_PER_IMAGE, "l2": "per-step working set 2.5 GB >> 126 MB L2 (no flush needed)",
                       "precision": "tf32", "peaks": peak_src},
            "images_per_sec": value / ITERS_PER_IMAGE,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": int(z0.numel() * 4),
                    "d2h_bytes_per_step": 4, "steps": e2e_steps, "last_loss": last,
                    "api": "models.get_net(...).type(cuda) + closure-style zero_grad/forward/MSELoss/backward/FusedAdam.step"},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "roofline": {"kernel": "tc_conv_kernel, dominant launch: level-0 3x3 conv 132->128 @512x512 fprop "
                                   "(tcgen05 tf32 implicit GEMM, smem input patch shared by the 9 taps)",
                         "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf32_peak if tf32_peak else None, "traffic": ROOFLINE_TRAFFIC_BYTES,
                         "algorithmic_flops_per_launch": dom_fl / max(dom_n, 1), "launches": dom_n,
                         "us_per_launch": 1000.0 * dom_ms / max(dom_n, 1),
                         "peak_note": "tf32 dense = 1/2 of the measured sustained bf16 cuBLAS rate (" + peak_src + ")",
                         "timed": "CUDA events around every launch in a %d-step eager pass right after the "
                                  "graph-replayed timed region (side stream off so that kernels run alone)" % roof_steps},
            "roofline_all_conv": {"kernel": "tc_conv_kernel, all %d fprop+dgrad launches of a step" % (conv_n // max(roof_steps, 1)),
                                  "achieved": conv_all, "peak": tf32_peak, "unit": "TFLOP/s", "frac": conv_all / tf32_peak,
                                  "ms_per_step": conv_ms / roof_steps,
                                  "share_of_step": (conv_ms / roof_steps) / (ms / args.steps) if ms > 0 else None},
            "roofline_wgrad": {"kernel": "tc_wgrad_kernel (tcgen05 tf32, MN-major operands, split-K), all launches",
                               "bound": "tensor", "achieved": wg_ach, "peak": tf32_peak, "unit": "TFLOP/s",
                               "frac": wg_ach / tf32_peak if tf32_peak else None, "launches": wg_n,
                               "ms_per_step": wg_ms / roof_steps,
                               "share_of_step": (wg_ms / roof_steps) / (ms / args.steps) if ms > 0 else None},
            "roofline_hbm": {"kernel": "k_bn_bwd_apply<plain>: BatchNorm+LeakyReLU backward (apply pass) behind the level-0 up conv, "
                                       "128 ch @512x512 (largest HBM-bound launch of the step)",
                             "bound": "hbm", "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                             "algorithmic_bytes_per_launch": hbm_big, "launches": len(hbm_sel),
                             "us_per_launch": 1000.0 * hbm_ms / max(len(hbm_sel), 1),
                             "traffic": ROOFLINE_HBM_TRAFFIC_BYTES,
                             "note": "algorithmic bytes = read raw + read gradient + write input gradient (3 x 134.2 MB); "
                                     "traffic = dram read+write of one launch from profiles/r01_ncu_bn_bwd_apply_l0.txt "
                                     "(part of the gradient is still in L2 from the producing conv)"},
            "step_tflops": ALG_GFLOP_PER_ITER / 1000.0 / (ms_max / args.steps / 1000.0),
            "per_rank": [{"psnr_gt": r[0], "final_loss": r[1], "it_per_s": r[2]} for r in recs],
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = best_cpu_threads(torch)
            its, s_per = cpu_iterations(torch, 5, 1, cores)
            line["cpu_baseline"] = {"value": its, "unit": "it/s", "cores": cores, "kind": "port",
                                    "sample": "5 iterations after 1 warm-up of the same 512x512 workload (oracle port of "
                                              "the reference's torch-CPU path), %.2f s/iter, %d threads = best of a probe "
                                              "over {8,16,32,64,%d}" % (s_per, cores, os.cpu_count() or 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
