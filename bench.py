#!/usr/bin/env python
"""Headline benchmark: point-evals/s (forward + loss + backward + Adam) of the NIF hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Workload (BASELINE.json configs[1], SURVEY 8 "C2"): NIFMultiScale, ShapeNet 2 -> 4x64 SIREN -> 1
(omega_0 30), ParameterNet 1 -> 64 x4 swish shortcut MLP -> latent 32, po_dim 16 897, 1 M synthetic
(t, x0, x1) points, fp32.  One step = one optimisation step over one batch of 65 536 points per GPU (the throughput batch
SURVEY 8(d) fixes for C2).  `--batch B` measures another batch: the forward and reverse kernels run one persistent CTA per SM
on pairs of 128-row tiles, so a multiple of 148 SMs x 256 rows = 37 888 fills their last wave -- 113 664 rows: 66.5-69.1 M
rows/s against 60.4-60.8 M at 65 536, whose last wave is 13 % empty (profiles/r02w_wave_batches.json, r02x_*, r02y_*).
N > 1 (torchrun, one rank per GPU): weak scaling, per-GPU batch fixed, one NCCL all-reduce of the flat
gradient buffer per step.

`--impl reference` times the reference's CPU implementation of the same step: TensorFlow cannot be
installed in this image, so it is the oracle port (oracle/nif_oracle.py, materialised (B, po_dim)
dataflow + autograd + TF-semantics Adam) on all host threads, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG_S = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 4,
         "weight_init_factor": 0.01, "omega_0": 30.0}
CFG_P = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}
N_POINTS = 1_000_000
BATCH = 65_536  # SURVEY 8(d); --batch overrides (113 664 = 148 SMs x 3 pairs of 128-row tiles fills the kernels' last wave)
METRIC = "point-evals/sec (fwd+bwd+Adam)"
WORKLOAD = ("C2 tutorial-2 multi-scale NIF: ShapeNet 2->4x64->1 SIREN (omega0 30), ParameterNet 1->64x4 swish->latent 32, "
            "po_dim 16897, 1M points/GPU, fp32")
UNIT = "points/s"


def synth_c2(n, seed=0):
    """t ~ U(-1,1), x ~ U(-1,1)^2, u = exp(-10 s^2) sin(40 s), s = x0 + 0.5 x1 - 0.3 t, max|u| = 1."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(-1, 1, (n, 1))
    x = rng.uniform(-1, 1, (n, 2))
    s = x[:, :1] + 0.5 * x[:, 1:2] - 0.3 * t
    u = np.exp(-10 * s**2) * np.sin(40 * s)
    u /= np.abs(u).max()
    return np.hstack([t, x]).astype(np.float32), u.astype(np.float32)


def flops_per_point():
    """F_fwd = 2 K P + 2 W_s + F_trunk; F_step = 3 F_fwd."""
    si, so, n, l, K = 2, 1, 64, 4, 32
    W_s = si * n + l * n * n + n * so
    P = W_s + (l + 1) * n + so
    n_st, l_st, pi = 64, 4, 1
    F_trunk = 2 * (pi * n_st + l_st * n_st * n_st + n_st * K)
    F_fwd = 2 * K * P + 2 * W_s + F_trunk
    return F_fwd, 3 * F_fwd, P


# ------------------------------------------------------------------------------------------------------
def cpu_reference_rate(rows_per_step, steps, warmup, threads=None):
    """points/s of the oracle port (materialised dataflow) on the host cores."""
    from oracle import nif_oracle as O
    if threads:
        torch.set_num_threads(threads)
    spec = O.spec_from_cfg("NIFMultiScale", CFG_S, CFG_P)
    tr = O.MaterialisedTrainer(spec, O.init_params(spec, 0), lr=1e-3)
    inp, tgt = synth_c2(rows_per_step * (steps + warmup), seed=1)
    inp, tgt = torch.as_tensor(inp), torch.as_tensor(tgt)
    for i in range(warmup):
        tr.step(inp[i * rows_per_step:(i + 1) * rows_per_step], tgt[i * rows_per_step:(i + 1) * rows_per_step])
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        tr.step(inp[i * rows_per_step:(i + 1) * rows_per_step], tgt[i * rows_per_step:(i + 1) * rows_per_step])
    dt = time.perf_counter() - t0
    return rows_per_step * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = 2048
    cores = os.cpu_count() or 1
    rate, sec = cpu_reference_rate(rows, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": rows,
                   "note": "TensorFlow 2.11 is not installable here; oracle port of the reference's materialised "
                           "(B,po_dim) dataflow, torch CPU fp32, on a bounded sample of the workload per step"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {rows} rows of the C2 workload (fwd+bwd+Adam)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [s.strip() for s in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def run_ours(args):
    import nif_b200
    from nif_b200.distributed import DataParallel

    dp = DataParallel("nccl")
    dev = dp.device
    torch.cuda.set_device(dev)
    rank, world = dp.rank, dp.world

    net = nif_b200.NIFMultiScale(CFG_S, CFG_P, "float32", seed=0, device=dev)
    model = net.build()
    model.compile(nif_b200.Adam(1e-3), loss="mse")
    if world > 1:
        dp.attach(model)

    # every rank owns a disjoint 1M-point shard (weak scaling); batches rotate through it
    inp_h, tgt_h = synth_c2(N_POINTS, seed=100 + rank)
    nb = N_POINTS // BATCH
    inp_pin = torch.as_tensor(inp_h[: nb * BATCH]).view(nb, BATCH, 3).pin_memory()
    tgt_pin = torch.as_tensor(tgt_h[: nb * BATCH]).view(nb, BATCH, 1).pin_memory()
    inp_d, tgt_d = inp_pin.to(dev), tgt_pin.to(dev)
    gb = BATCH * world

    def step_resident(i):
        return model._train_step(inp_d[i % nb], tgt_d[i % nb], None, gb)

    def step_e2e(i):
        # the public call: pinned HOST batch in (copied to the device inside), the step's loss read back as a float
        return model.train_on_batch(inp_pin[i % nb], tgt_pin[i % nb], global_batch=gb)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        dp.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        dp.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dp.max_(ms)
        return float(ms) / steps

    sampler = ClockSampler(dev.index if dev.index is not None else 0) if rank == 0 else None
    ms_step = timed(step_resident, args.steps, args.warmup)
    ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup // 2))
    clocks = sampler.stop() if sampler else None  # sampled under load across both timed regions

    # ---- sustained rate: >= 3 s of back-to-back steps after the timed regions, with its own clock sample ----
    sampler2 = ClockSampler(dev.index if dev.index is not None else 0) if rank == 0 else None
    n_sus = max(args.steps, int(3000.0 / max(ms_step, 1e-3)) + 1)
    ms_sus = timed(step_resident, n_sus, 3)
    clocks_sus = sampler2.stop() if sampler2 else None

    # data-parallel update path: NVSwitch multicast fused kernel, or NCCL all-reduce + Adam on every rank
    update_mode = "single" if world == 1 else ("multimem reduce-scatter+Adam+all-gather kernel" if model._symm is not None
                                               else "NCCL all-reduce + Adam")
    # ---- replicas agree: after all those steps every rank must hold bit-identical parameters ----
    params_equal = None
    if world > 1:
        lo, hi = net.theta.clone(), net.theta.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        params_equal = bool(torch.equal(lo, hi))

    if rank != 0:
        dp.shutdown()
        return

    # ---- per-kernel table (outside the timed regions): eager launches, every launch of the library bracketed by CUDA
    # events on its own stream (nif_profile_begin / nif_profile_end); isolates which kernel dominates the step ----
    from nif_b200.ops import kernel_profile
    eng = net.engine
    F_fwd, F_step, P = flops_per_point()
    model.use_graph = False
    model.dist = None  # the table is this rank's kernels; the collective is not a kernel of the library
    for i in range(2):
        step_resident(i)
    torch.cuda.synchronize()
    n_prof = 5
    with kernel_profile() as prof:
        for i in range(n_prof):
            step_resident(i)
    W_s = 2 * 64 + 4 * 64 * 64 + 64
    K, H, n_s = 32, 4, 64
    alg_flops = {  # ALGORITHMIC flops per launch: rows x per-row figure
        "nif_tc_fwd_kernel": (2 * K * P + 2 * W_s) * BATCH,            # latent->weights projection + ShapeNet, forward
        "nif_tc_bwd_data_kernel": (2 * K * P + 2 * W_s) * BATCH,       # the same products, transposed weights
        "nif_tc_bwd_weight_kernel": 2 * (K + 1) * H * n_s * n_s * BATCH,  # batch reduction of zt (x) h (x) da, hidden matrices
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_sus = float(peaks.get("bf16_tflops_sustained", 1590.0 * 0.88))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic_by_kernel, traffic_file = {}, None
    # DRAM bytes per launch from the committed `ncu --set full` capture of one step AT THIS BATCH (the captures record
    # their batch: r02_ncu_kernels.json is the 65 536-row one, r02x_ncu_kernels.json the 113 664-row one)
    for fn, fb in (("r02_ncu_kernels.json", 65536), ("r02x_ncu_kernels.json", None)):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", fn)))
        except (OSError, ValueError):
            continue
        if int(t.pop("_batch", fb or 0)) == BATCH:
            traffic_by_kernel, traffic_file = t, fn
            break
    tot_ms = sum(ms for _, _, ms in prof.table) or 1.0
    table = []
    for name, cnt, ms in sorted(prof.table, key=lambda r: -r[2]):
        us = ms * 1e3 / cnt
        row = {"kernel": name, "launches_per_step": cnt / n_prof, "us_per_launch": us, "share_of_step": ms / tot_ms}
        if name in alg_flops:
            row["tflops"] = alg_flops[name] / (us * 1e-6) / 1e12
            row["frac_of_burst_peak"] = row["tflops"] / peak_burst
        if name in traffic_by_kernel:
            row["dram_bytes_per_launch"] = traffic_by_kernel[name]
        table.append(row)
    dom = table[0]
    launches_per_step = sum(cnt for _, cnt, _ in prof.table) / n_prof
    fp32_peak = nif_b200.ops.measure_fp32_peak()
    step_tflops = F_step * BATCH / (ms_step * 1e-3) / 1e12
    # algorithmic HBM bytes per step: inputs + targets + Adam (28 B/param) + gradient write/read (8 B/param)
    n_par = net.count_params()
    hbm_bytes = BATCH * 4 * (1 + 2 + 1) + 36 * n_par
    pts = BATCH * world / (ms_step * 1e-3)
    pts_e2e = BATCH * world / (ms_e2e * 1e-3)

    cores = os.cpu_count() or 1
    cpu_rate, cpu_sec = cpu_reference_rate(2048, 6, 1, cores) if world == 1 else (None, None)

    line = {
        "metric": METRIC, "value": pts, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": gb, "parallelism": f"dp{world}",
                   "kernels": eng.kernel_path, "trunk_kernels": getattr(net._trunk, "kernel_path", None),
                   "dp_update": update_mode,
                   "batch_note": ("65 536 rows: SURVEY 8(d)'s throughput batch for C2 (256 tile pairs on 148 SMs: the last "
                                  "wave of the persistent forward / reverse kernels is 13 % empty; --batch 113664 = 148 x 3 "
                                  "pairs fills it: profiles/r02x_bench_1gpu.json, r02y_bench_*.json)") if BATCH == 65536 else
                                 f"--batch {BATCH} (the default is SURVEY 8(d)'s 65 536 rows)",
                   "l2": "inputs rotate through the 1M-point set; per-step working set (activation stash + deltas, "
                         "> 200 MB) exceeds the 126 MB L2, no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": pts_e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": BATCH * 4 * 4,
                "d2h_bytes_per_step": 4},
        "sustained": {"value": BATCH * world / (ms_sus * 1e-3), "unit": UNIT, "ms_per_step": ms_sus, "steps": n_sus,
                      "seconds": ms_sus * n_sus / 1e3, "clocks": clocks_sus},
        "params_equal_across_ranks": params_equal,
        "gpu_launches": int(round(launches_per_step * args.steps)),
        # the DOMINANT kernel of the step by time (kernel table below), timed by CUDA events around its launches inside
        # eager steps, against the burst dense 16-bit peak
        "roofline": {"bound": "tensor", "kernel": dom["kernel"], "achieved": dom.get("tflops"),
                     "peak": peak_burst, "unit": "TFLOP/s",
                     "frac": (dom["tflops"] / peak_burst) if "tflops" in dom else None,
                     "traffic": dom.get("dram_bytes_per_launch"),
                     "traffic_note": (f"dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                      f"(profiles/{traffic_file})") if traffic_file else
                                     "no ncu --set full capture at this batch under profiles/",
                     "peak_source": f"bf16_tflops (burst, dense 16-bit MMA), {peak_src}",
                     "note": "achieved = ALGORITHMIC flops per launch / launch time; the fp32-grade FP16x3 split issues 3 "
                             "tensor-core MACs per algorithmic MAC, so 1/3 of peak is this path's ceiling",
                     "tensor_macs_per_algorithmic_mac": 3 if eng.kernel_path == "fp16x3" else 1,
                     "us_per_launch": dom["us_per_launch"], "share_of_step": dom["share_of_step"],
                     "kernel_table": table,
                     "whole_step": {"tflops": step_tflops, "peak": peak_sus, "frac": step_tflops / peak_sus,
                                    "peak_source": f"bf16_tflops_sustained, {peak_src}"},
                     "fp32_fma_peak_tflops": fp32_peak,
                     "hbm": {"achieved_gbs": hbm_bytes / (ms_step * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "frac": hbm_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak,
                             "note": "compute-bound by design: 16 B/point + 36 B/param per step"}},
        "cpu_baseline": None if cpu_rate is None else {
            "value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "6 steps x 2048 rows of the C2 workload, oracle port (materialised dataflow, torch CPU fp32)"},
    }
    print(json.dumps(line), flush=True)
    dp.shutdown()


def main():
    global BATCH
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="rows per step and GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.batch < 1 or args.batch > N_POINTS:
        raise SystemExit(f"--batch must be in 1..{N_POINTS}")
    BATCH = args.batch
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a GPU (nif_b200 has no CPU path); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
