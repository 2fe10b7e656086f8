#!/usr/bin/env python
"""bench.py -- images/sec of SDVAR draft-then-verify generation (d16 draft -> d30 target, 256 px) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one call of ``SDVAR.sdvar_autoregressive_infer_cfg_parallel_v1`` on a batch of B=64 synthetic class labels
per GPU (BASELINE.json configs[1]); data-parallel replicas, weak scaling (per-GPU batch fixed), NCCL only to gather the
images and the acceptance counters.  One JSON line is printed by rank 0:
  value      images/s, labels already resident in HBM, result left in HBM          (device-timed, max over ranks)
  e2e        same metric through the public API with HOST labels (pinned) -> device and the images read back to host
  roofline   the dominant kernel family (tcgen05 GEMM) : algorithmic FLOPs / CUDA-event time, vs the measured cuBLAS peak
  kernels    the same figure for every kernel family (GB/s for the HBM-bound ones), incl. the verify kernel of the metric
  cpu_baseline  the oracle restatement of the reference's loop timed on this box's host cores on a bounded sample
``--impl reference`` times that CPU implementation alone (the reference is pure Python/PyTorch and does not travel to the
GPU box; the oracle port under oracle/ is what is timed -- "kind": "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
P512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)
METRIC = "images/sec (d16->d30 SD, 256px)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--depth-draft", type=int, default=16)
    ap.add_argument("--depth-target", type=int, default=30)
    ap.add_argument("--gamma", type=int, default=2)
    ap.add_argument("--cfg", type=float, default=1.5)
    ap.add_argument("--top-k", type=int, default=900)
    ap.add_argument("--top-p", type=float, default=0.96)
    ap.add_argument("--accept-rule", default="speculative", choices=["speculative", "reference"])
    ap.add_argument("--px", type=int, default=256, choices=[256, 512], help="256: patch_nums 1..16 (L=680); 512: 1..32 (L=2240)")
    ap.add_argument("--shared-aln-target", action="store_true", help="target uses shared adaLN (the d36 layout, README.md:142-144)")
    ap.add_argument("--cpu-sample-images", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiler runs only; such a line is not a bench value)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """samples nvidia-smi during the timed region (B200_PROFILING.md 'clocks line')"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(args, n_images: int, steps: int, warmup: int, device_for_init):
    """The oracle restatement of the reference loop (oracle/ref_model.py:sd_generate + decoder) on the host cores."""
    from oracle.ref_model import RefDecoder, RefVAR, RefVQ, ReplayNoise, sd_generate
    from sdvar_b200.weights import var_state_dict, vqvae_state_dict
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    cpu = lambda sd: {k: v.cpu() for k, v in sd.items()}
    vsd = cpu(vqvae_state_dict(ch=160, patch_nums=P256, device=device_for_init))
    d = RefVAR(cpu(var_state_dict(args.depth_draft, patch_nums=P256, seed=1, tag="draft", device=device_for_init)), P256)
    t = RefVAR(cpu(var_state_dict(args.depth_target, patch_nums=P256, seed=2, tag="target", device=device_for_init,
                                  shared_aln=args.shared_aln_target)), P256)
    vq, dec = RefVQ(vsd, P256), RefDecoder(vsd)
    lab = torch.randint(0, 1000, (n_images,), generator=torch.Generator().manual_seed(0))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        f_hat, _, stats = sd_generate(d, t, vq, n_images, lab, ReplayNoise(i), cfg=args.cfg, gamma=args.gamma, top_k=args.top_k,
                                      top_p=args.top_p, accept_rule=args.accept_rule)
        dec.fhat_to_img(f_hat).add_(1).mul_(0.5)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return dict(value=n_images / dt, unit="images/s", cores=cores, kind="port",
                sample=f"{n_images} image(s)/step x {steps} step(s), fp32, d{args.depth_draft}->d{args.depth_target} SD gamma={args.gamma} "
                       f"oracle/ref_model.py:sd_generate + decoder, {warmup} warm-up"), dt, stats


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner on the first collective) must not share stdout with the ONE JSON line:
    fd 1 points at stderr while the benchmark runs and is restored by _emit."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)
    if _REAL_STDOUT is not None:
        os.dup2(2, 1)


def main():
    args = parse()
    _quiet_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    has_cuda = torch.cuda.is_available()
    global P256
    if args.px == 512:
        P256 = P512          # every use below takes the selected pyramid
    workload = (f"SDVAR VAR-d{args.depth_draft} draft + VAR-d{args.depth_target} target, random-init, {args.px}px, patch_nums 1..{P256[-1]}, "
                f"batch {args.batch}/GPU, cfg={args.cfg}, top_k={args.top_k}, top_p={args.top_p}, gamma={args.gamma}, accept_rule={args.accept_rule}")

    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = f"cuda:{local}" if has_cuda else "cpu"
        cb, dt, stats = cpu_reference_run(args, args.cpu_sample_images, args.steps, args.warmup, dev)
        _emit({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                          "config": {"workload": workload, "note": "CPU reference arm: bounded sample of the same workload"},
                          "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0, "accept_stats": {k: stats[k] for k in ("rounds", "target_passes", "accepted_tokens", "rejected_tokens")}})
        return 0

    if not has_cuda:
        _emit({"error": "no CUDA device: sdvar_b200 has no CPU fallback"})
        return 2
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from sdvar_b200 import _cabi, parallel
    from sdvar_b200.models import build_vae_var_speculative_decoding
    from sdvar_b200.weights import var_state_dict, vqvae_state_dict

    torch.manual_seed(0)
    vae, draft, target, sd = build_vae_var_speculative_decoding(dev, patch_nums=P256, depth_draft=args.depth_draft, depth_target=args.depth_target,
                                                                shared_aln_target=args.shared_aln_target)
    vae.load_state_dict(vqvae_state_dict(ch=160, patch_nums=P256, device=dev))
    draft.load_state_dict(var_state_dict(args.depth_draft, patch_nums=P256, seed=1, tag="draft", device=dev))
    target.load_state_dict(var_state_dict(args.depth_target, patch_nums=P256, seed=2, tag="target", device=dev, shared_aln=args.shared_aln_target))
    B = args.batch
    px = 16 * P256[-1]
    lab_host = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(rank)).pin_memory()
    lab_dev = lab_host.to(dev)
    img_host = torch.empty(B, 3, px, px, dtype=torch.float32).pin_memory()

    def step(i: int, e2e: bool):
        lab = lab_host.to(dev, non_blocking=True) if e2e else lab_dev
        img = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, g_seed=1000 * rank + i, cfg=args.cfg, gamma=args.gamma,
                                                           top_k=args.top_k, top_p=args.top_p, accept_rule=args.accept_rule)
        st = sd.last_stats
        if world > 1:   # the path's only collectives: images + acceptance counters
            parallel.gather_images(img)
            parallel.reduce_stats(st, dev)
        if e2e:
            img_host.copy_(img, non_blocking=True)
        return st

    def timed(e2e: bool):
        for i in range(args.warmup):
            step(i, e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local)
        sampler.start()
        torch.cuda.profiler.start()      # ncu --profile-from-start off captures exactly the timed region
        l0 = _cabi.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        last = None
        for i in range(args.steps):
            last = step(args.warmup + i, e2e)
        b.record()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        ms = a.elapsed_time(b)
        launches = _cabi.launch_count() - l0
        clocks = sampler.stop()
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms, launches, clocks, last

    ms_dev, launches, clocks, last_stats = timed(False)
    ms_e2e = timed(True)[0] if not args.no_e2e else float("nan")
    total_imgs = world * B * args.steps
    value = total_imgs / (ms_dev * 1e-3)
    e2e_value = total_imgs / (ms_e2e * 1e-3)

    pk = peaks()
    out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic",
           "config": {"workload": workload, "global_batch": world * B, "parallelism": f"dp{world}",
                      "l2": "no flush: per-step working set (bf16 weights 4.6 GB + KV ring ~27 GB + logits) >> 126 MB L2",
                      "weights": "sdvar_b200.weights hashed init (seed 1 draft / 2 target / 0 vae)", "labels": "randint(0,1000) seed=rank"},
           "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": B * 8,
                   "d2h_bytes_per_step": B * 3 * px * px * 4},
           "gpu_launches": launches, "clocks": clocks,
           "accept_stats": {k: last_stats[k] for k in ("rounds", "target_passes", "draft_stages", "accepted_tokens", "rejected_tokens", "advance")}}

    if not args.no_profile:   # roofline leg: one extra step with per-family CUDA-event timing on the launch stream
        _cabi.profile_begin()
        st = step(10_000, False)
        prof = _cabi.profile_end()
        kern = {}
        for fam, (ms, work, n) in prof.items():
            if n == 0:
                continue
            if fam in ("gemm", "attention"):
                kern[fam] = {"bound": "tensor", "achieved": work / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms": ms, "launches": n}
            else:
                extra = st["rejected_tokens"] * 4096 * 4.0 if fam == "verify" else 0.0   # resample noise is read only on reject
                kern[fam] = {"bound": "hbm", "achieved": (work + extra) / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms, "launches": n}
        g = kern.get("gemm")
        if g:
            # DRAM bytes per launch of the GEMM family from the committed ncu launch list of this same default workload
            # (profiles/traffic_r01.json <- profiles/launches_r01.md); null for any other workload
            traffic = None
            tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r01.json")
            default_workload = (args.batch, args.depth_draft, args.depth_target, args.gamma, args.px, args.top_k) == (64, 16, 30, 2, 256, 900)
            if default_workload and os.path.exists(tpath):
                traffic = json.load(open(tpath))["families"]["gemm"]["dram_bytes_per_launch"]
            out["roofline"] = {"bound": "tensor", "kernel": "sdvar::gemm2::gemm2_kernel<EPI> / gemm::gemm_kernel<EPI> (tcgen05, all epilogues)", "achieved": g["achieved"],
                               "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": g["achieved"] / pk["tf_sustained"], "traffic": traffic,
                               "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the family's launches)",
                               "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                               "share_of_step_ms": g["ms"], "launches": g["launches"]}
        for k, v in kern.items():
            v["frac"] = v["achieved"] / (pk["tf_sustained"] if v["bound"] == "tensor" else pk["hbm"])
        out["kernels"] = kern

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _, _ = cpu_reference_run(args, args.cpu_sample_images, 1, 0, dev)
        out["cpu_baseline"] = cb
    if rank == 0:
        _emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
