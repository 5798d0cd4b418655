#!/usr/bin/env python
"""bench.py — diffusion timesteps/s of the sampler hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference algorithm's CPU port (oracle/) on the host cores

A "step" is one diffusion timestep of the whole batch: score-network forward -> IGSO(3)/R^3 scores -> reverse SDE
step -> backbone atoms of x_{t-1} and of the x0 prediction (experiments/utils.py:292-412 of the reference).
Workload at N=1: BASELINE.json configs[1] (TCR CDR3 inpainting, N_res=350, batch 8, 500-timestep schedule); with
N>1 every rank runs its own batch of 8 independent samples (weak scaling; samples shard with no data-path collective;
NCCL is used for the weight broadcast before and the gather of final coordinates after the loop).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "diffusion timesteps/sec (batch x N_res)"
UNIT = "residue-timesteps/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 0) or 0),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 0) or 0), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU DURING the timed region from a background thread through NVML
    (nvidia_ml_py); polling `nvidia-smi -lms` from a subprocess instead was measured to stall kernel launches of the
    timed region (driver lock), so it is only the fallback when NVML cannot be imported."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int, period_s: float = 0.02):
        import threading

        self.sm, self.mem, self.reasons, self.sm_max = [], [], set(), None
        self._armed = False
        self._stop = threading.Event()
        self.p = self.f = self.th = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}

            def loop():
                while not self._stop.is_set():
                    if os.environ.get("FDPT_BENCH_NO_POLL") and self._armed:
                        self._stop.wait(period_s)
                        continue
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mem.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    self._stop.wait(period_s)

            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
            self._nvml_ok = True
        except Exception:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu)],
                                          stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def begin(self):
        """Forget what was sampled so far: called right before the timed region (the sampler itself is created -- NVML initialised --
        before the warm-up, so that no idle gap separates warm-up and timed region)."""
        self.sm.clear()
        self.mem.clear()
        self.reasons.clear()
        self._armed = True

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": []}
        if self.th is not None:
            self._stop.set()
            self.th.join(timeout=2)
            if self.sm:
                out["sm_mhz"] = float(np.median(self.sm))
                out["sm_min_mhz"] = float(np.min(self.sm))
                if self.mem:
                    out["mem_min_mhz"] = float(np.min(self.mem))
                out["samples"] = len(self.sm)
            out["reasons"] = sorted(self.reasons)
            out["source"] = "nvml"
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        out["source"] = "nvidia-smi"
        return out


def cpu_port_rate(wl, steps: int, warmup: int, threads: int):
    """Times the oracle port (CPU restatement of the reference) on a bounded sample: B=1 of the workload."""
    import torch
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.config import default_conf
    from framedipt_b200.params import synthetic_state_dict
    from oracle import framedipt_oracle as orc

    torch.set_num_threads(threads)
    conf = default_conf()
    diffuser = SE3Diffuser(conf.diffuser)
    sd = synthetic_state_dict(0, with_aatype=not wl.de_novo)
    np.random.seed(123)
    feats = synthetic.make_features(wl, diffuser, seed=0, batch=1)
    n = wl.n_res
    # one self-conditioning-free trajectory segment: (warmup + steps) timesteps of the real schedule
    sched_t = np.linspace(wl.min_t, 1.0, wl.num_t)[::-1]
    dt = 1 / wl.num_t
    noise = np.random.normal(size=(warmup + steps, 2, 1, n, 3))
    dm = ((1 - feats["fixed_mask"]) * feats["res_mask"]).numpy().astype(np.float64)
    aat = orc.preprocess_aatype(feats.get("aatype"), feats["fixed_mask"], not wl.de_novo, not wl.de_novo)
    t_start = None
    with torch.no_grad():
        for s in range(warmup + steps):
            if s == warmup:
                t_start = time.perf_counter()
            t = sched_t[s]
            feats["t"] = t * torch.ones(1)
            out = orc.score_network_forward(sd, feats, inpainting=not wl.de_novo, input_aatype=not wl.de_novo)
            rig = out["rigids"].float()
            feats["sc_ca_t"] = rig[..., 4:]
            R1, T1 = orc.reverse_step(feats["rigids_t"].float().numpy(), out["rot_score"].numpy().astype(np.float64),
                                      out["trans_score"].float().numpy(), dm, t, dt, noise[s, 0], noise[s, 1], noise_scale=wl.noise_scale)
            q1 = orc.rot_to_quat_np(R1.astype(np.float64)).astype(np.float32)
            feats["rigids_t"] = torch.tensor(np.concatenate([q1, T1], -1))
            orc.compute_backbone(orc.quat_to_rot(rig[..., :4]), rig[..., 4:], out["psi"].float(), aat)
            orc.compute_backbone(R1, T1, out["psi"].float(), aat)
    el = time.perf_counter() - t_start
    return n * steps / el, el / steps


def eager_gpu_rate(wl, B: int, steps: int, warmup: int, dev):
    """Secondary baseline (recommended): what a user of the reference gets on this box with `use_gpu: true` -- the same
    eager PyTorch modules on the B200 (fp32, TF32 off), state round-tripping to the host every step for the numpy reverse step
    (experiments/utils.py:292-412).  The reference checkout does not exist on the GPU box, so its restatement (oracle port: the same
    torch ops in the same order) is what runs; factory calls inside it are routed to the GPU with torch.device(...)."""
    import torch
    from framedipt_b200 import SE3Diffuser, synthetic
    from framedipt_b200.config import default_conf
    from framedipt_b200.params import synthetic_state_dict
    from oracle import framedipt_oracle as orc

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    diffuser = SE3Diffuser(default_conf().diffuser)
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(0, with_aatype=not wl.de_novo).items()}
    np.random.seed(123)
    feats = {k: v.to(dev) for k, v in synthetic.make_features(wl, diffuser, seed=0, batch=B).items()}
    n = wl.n_res
    sched_t = np.linspace(wl.min_t, 1.0, wl.num_t)[::-1]
    dt = 1 / wl.num_t
    noise = np.random.normal(size=(warmup + steps, 2, B, n, 3))
    dm = ((1 - feats["fixed_mask"]) * feats["res_mask"]).cpu().numpy().astype(np.float64)
    t_start = None
    with torch.no_grad(), torch.device(dev):
        for s in range(warmup + steps):
            if s == warmup:
                if torch.device(dev).type == "cuda":
                    torch.cuda.synchronize(dev)
                t_start = time.perf_counter()
            t = float(sched_t[s])
            feats["t"] = t * torch.ones(B)
            out = orc.score_network_forward(sd, feats, inpainting=not wl.de_novo, input_aatype=not wl.de_novo)
            rig = out["rigids"].float()
            feats["sc_ca_t"] = rig[..., 4:]
            with torch.device("cpu"):  # the reverse step is host numpy in the reference (its torch helpers must not land on the GPU)
                R1, T1 = orc.reverse_step(feats["rigids_t"].float().cpu().numpy(), out["rot_score"].cpu().numpy().astype(np.float64),
                                          out["trans_score"].float().cpu().numpy(), dm, t, dt, noise[s, 0], noise[s, 1],
                                          noise_scale=wl.noise_scale)
                q1 = orc.rot_to_quat_np(R1.astype(np.float64)).astype(np.float32)
                nxt = torch.tensor(np.concatenate([q1, T1], -1))
            feats["rigids_t"] = nxt.to(dev)
        if torch.device(dev).type == "cuda":
            torch.cuda.synchronize(dev)
    el = time.perf_counter() - t_start
    return B * n * steps / el, el / steps


DTYPE = "mixed: fp16-operand/fp32-accumulate tcgen05 on the pair side (z stored fp16), 2-term fp16 split (fp32-class) on the node side, fp32 SIMT elsewhere, fp64 SDE step"


def run_workload(ctx, model, diffuser, wl, B, K, W, dev, dist, world, spinup=True, seed_rank=0):
    """Times K consecutive timesteps of workload `wl` with B samples on this rank (state resident in HBM); returns
    (ms over the K steps on this rank, launches in the timed region, diag dict, prepared feats, schedule segment fn, noise tensors)."""
    import torch

    from framedipt_b200 import synthetic
    from framedipt_b200.inference import build_schedule

    np.random.seed(123 + seed_rank)
    feats = synthetic.make_features(wl, diffuser, seed=0, batch=B)
    N = wl.n_res
    _, sched_full, temb_full = build_schedule(diffuser, wl.num_t, wl.min_t, wl.noise_scale)
    assert K + W < wl.num_t

    def segment(s0, n):  # n consecutive timesteps of the real schedule, all doing a reverse step
        sc = sched_full[s0:s0 + n].copy()
        sc[:, 7] = 0.0
        return sc, temb_full[s0:s0 + n].contiguous()

    feats_dev = {k: v.to(dev) for k, v in feats.items()}
    pf = model.prepare(feats_dev, dev)
    noise_host = torch.from_numpy(np.random.normal(size=(K + W, 2, B, N, 3))).pin_memory()
    noise_dev = noise_host.to(dev)
    ctx.reserve(B, N)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # clock spin-up (untimed, not counted as warm-up): a fresh box idles in a low power state; the load is the workload itself
    sc, te = segment(0, W)
    t_spin = time.perf_counter()
    while spinup and time.perf_counter() - t_spin < 0.6:
        ctx.sample(pf, sc, te, noise_dev[:W], self_condition=False)
        torch.cuda.synchronize(dev)
    # warm-up (untimed)
    out = ctx.sample(pf, sc, te, noise_dev[:W], self_condition=True)
    pf.rigids_t = out["rigid_traj"][0].contiguous()
    pf.sc_ca_t = out["trans_traj"][0].contiguous()
    torch.cuda.synchronize(dev)
    # timed region
    sc, te = segment(W, K)
    te = te.to(dev)
    out_buf = ctx.alloc_traj(B, N, K)  # allocated before the clock starts (a cold cudaMalloc costs 1 - 60 ms)
    l0 = ctx.launch_count()
    barrier()
    return dict(pf=pf, sc=sc, te=te, noise_dev=noise_dev, noise_host=noise_host, out_buf=out_buf, l0=l0, barrier=barrier, segment=segment, feats=feats_dev)


def time_steps(ctx, st, K, W, dev, dist, clocks=None):
    import torch

    if clocks:
        clocks.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cap0, wall0 = ctx.stat(0), time.perf_counter()
    ev0.record()
    ctx.sample(st["pf"], st["sc"], st["te"], st["noise_dev"][W:], self_condition=False, out=st["out_buf"])
    ev1.record()
    wall_enqueue = time.perf_counter() - wall0
    st["barrier"]()
    ms = ev0.elapsed_time(ev1)
    diag = {"graph_captures_in_timed_region": ctx.stat(0) - cap0, "enqueue_wall_ms": round(wall_enqueue * 1e3, 3),
            "sample_host_ms": {k: round(v, 3) for k, v in ctx.last_sample_host_ms.items()}}
    launches = ctx.launch_count() - st["l0"]
    t_all = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    return float(t_all.item()), int(launches), diag


def quick_rate(ctx, model, diffuser, wl, B, K, W, dev, dist, world):
    """ms per step (max over ranks) of a secondary workload: the same timed-region protocol, no e2e / profile legs."""
    st = run_workload(ctx, model, diffuser, wl, B, K, W, dev, dist, world, spinup=False)
    ms, _, _ = time_steps(ctx, st, K, W, dev, dist)
    return ms / K


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--debug-flags", type=int, default=0, help="FDPT_OPT_DEBUG_FLAGS bit set (A/B switches of include/fdpt.h)")
    ap.add_argument("--no-spinup", action="store_true", help="skip the untimed clock spin-up (for ncu launch lists, where every launch is expensive)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_tcr350")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: shard this many samples over the ranks (sharding.shard_range) instead of wl.batch samples per rank")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the CPU baseline leg (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary-workload lines (cfg3 at N=1, cfg4 at N=4, cfg5 at N=8)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: W >= 3"

    from framedipt_b200 import sharding, synthetic

    wl = synthetic.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    if args.global_batch:
        b0, b1 = sharding.shard_range(args.global_batch, rank, world)
        B = b1 - b0
        scaling, total_B = "strong", args.global_batch
    else:
        B, scaling, total_B = wl.batch, "weak", wl.batch * world
    config = {"workload": f"{wl.name}: B={B}/GPU ({total_B} in all), N_res={wl.n_res}, schedule num_t={wl.num_t}, inpainting={not wl.de_novo}",
              "batch_per_gpu": B, "global_batch": total_B, "n_res": wl.n_res, "parallelism": f"sample-parallel x{args.gpus}",
              "l2_policy": f"inputs larger than L2: pair representation z (fp16) = {B * wl.n_res ** 2 * 256 / 1e6:.0f} MB, streamed 11x per forward"}

    if args.impl == "reference":
        if rank != 0:
            return
        rate, sps = cpu_port_rate(wl, args.steps, args.warmup, cores)
        config = dict(config, batch_per_gpu=1, global_batch=1,
                      workload=f"{wl.name}: B=1 (the reference runs one sample at a time, experiments/sampler.py:352), N_res={wl.n_res}, inpainting={not wl.de_novo}")
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (torch CPU), f64 SDE step", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"oracle port (CPU restatement of the reference, pinned to reference fixtures), B=1, N_res={wl.n_res}, "
                                           f"{args.steps} timesteps after {args.warmup} warm-up; ms_per_step is the measured time of one B=1 timestep"},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf
    from framedipt_b200.inference import inference_fn
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.score_network import ScoreNetwork

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def build_model(de_novo):
        conf = default_conf(input_aatype=not de_novo)
        diffuser = SE3Diffuser(conf.diffuser)
        model = ScoreNetwork(conf.model, diffuser, inpainting=not de_novo)
        # weights: rank 0 owns them, NCCL broadcast of the packed blob over NVLink (C1 in SURVEY §7)
        sd = synthetic_state_dict(0, with_aatype=not de_novo)
        if world > 1:
            if rank != 0:
                sd = {k: torch.zeros_like(v) for k, v in sd.items()}
            sd = sharding.broadcast_state_dict(sd, dist, 0, dev)
        model.load_state_dict(sd)
        model = model.to(dev).eval()
        return model, diffuser, model.context(dev)

    model, diffuser, ctx = build_model(wl.de_novo)
    N = wl.n_res
    K, W = args.steps, args.warmup
    if args.debug_flags:
        ctx.set_option(3, args.debug_flags)
    clocks = ClockSampler(local_rank) if rank == 0 else None  # NVML initialisation happens here, before any GPU work is timed
    st = run_workload(ctx, model, diffuser, wl, B, K, W, dev, dist, world, spinup=not args.no_spinup, seed_rank=rank)
    ms_max, launches, diag = time_steps(ctx, st, K, W, dev, dist, clocks)
    clk = clocks.stop() if clocks else None
    value = total_B * N * K / (ms_max * 1e-3)
    pf, noise_dev, barrier = st["pf"], st["noise_dev"], st["barrier"]

    # ---------------- per-kernel timings: the same K steps once more with CUDA-event pairs around the hot kernels ----------------
    # (the event pairs sit between kernels, so this pass enqueues the step directly instead of replaying its CUDA graph)
    ctx.profile_enable(True)
    ctx.sample(pf, st["sc"], st["te"], noise_dev[W:], self_condition=False)
    torch.cuda.synchronize(dev)
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---------------- e2e: the call a FrameDiPT user makes -- inference_fn(model, diffuser, feats, num_t, ...) for the FULL schedule with
    # HOST inputs and HOST outputs: feature dict uploaded from pinned host memory, schedule build, the self-conditioning pre-pass, all
    # num_t timesteps, and the [T,B,N,37,3] / [T+1,B,N,7] / [T,B,N,3] trajectories materialised as numpy arrays (aux_traj=True, like
    # experiments/inference.py).  Headline mode: device RNG (Philox); the reference-stream mode (host numpy draws, uploaded) is reported
    # next to it.  Pass 0 of each is an untimed warm-up of exactly the same call.
    e2e = None
    if not args.no_e2e:
        feats_host = {k: v.cpu().pin_memory() for k, v in st["feats"].items()}
        e2e = {}
        for mode in ("philox", "numpy"):
            for e2e_pass in range(2):
                barrier()
                t0 = time.perf_counter()
                f_dev = {k: v.to(dev, non_blocking=True) for k, v in feats_host.items()}
                o = inference_fn(model, diffuser, f_dev, num_t=wl.num_t, min_t=wl.min_t, aux_traj=True, noise_scale=wl.noise_scale,
                                 inpainting=not wl.de_novo, input_aatype=not wl.de_novo, rng=mode, philox_seed=1234)
                torch.cuda.synchronize(dev)
                e2e_s = time.perf_counter() - t0
                if mode == "numpy":
                    break  # one pass: the draw of the full noise stream dominates its overhead, there is nothing to warm
            t_all = torch.tensor([e2e_s], device=dev)
            if dist is not None:
                dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
            e2e[mode] = float(t_all.item())
        assert o["prot_traj"].shape == (wl.num_t, B, N, 37, 3) and isinstance(o["prot_traj"], np.ndarray)
        if dist is not None:
            # final gather of the finished samples' coordinates (C2 in SURVEY §7), outside the step loop
            gathered = sharding.gather_samples(torch.from_numpy(o["prot_traj"][0][:, :, :5].copy()).to(dev), dist, 0)
            assert rank != 0 or gathered.shape[0] == total_B
    T_full = wl.num_t
    h2d_feat = sum(v.numel() * v.element_size() for v in st["feats"].values())
    d2h_step = int(B * N * (15 + 15 + 7 + 3) * 4)

    # ---------------- secondary workloads: the other BASELINE configs as they are stated ----------------
    extra = {}
    if not args.no_extra and args.workload == "cfg2_tcr350" and not args.global_batch:
        try:
            if world == 1:  # cfg3 = the largest single-GPU configuration
                w3 = synthetic.WORKLOADS["cfg3_denovo256"]
                m3, d3, c3 = build_model(True)
                ms3 = quick_rate(c3, m3, d3, w3, w3.batch, 5, 3, dev, None, 1)
                extra["cfg3_denovo256"] = {"batch_per_gpu": w3.batch, "n_res": w3.n_res, "ms_per_step": ms3,
                                           "value": w3.batch * w3.n_res / (ms3 * 1e-3), "unit": UNIT, "scaling": "single GPU"}
                del m3, c3
            if world == 4:  # cfg4: global batch 32 over 4 GPUs = 8 per GPU
                w4 = synthetic.WORKLOADS["cfg4_tcrpmhc800"]
                b0, b1 = sharding.shard_range(w4.batch, rank, world)
                ms4 = quick_rate(ctx, model, diffuser, w4, b1 - b0, 5, 3, dev, dist, world)
                extra["cfg4_tcrpmhc800"] = {"global_batch": w4.batch, "batch_per_gpu": b1 - b0, "n_res": w4.n_res, "ms_per_step": ms4,
                                            "value": w4.batch * w4.n_res / (ms4 * 1e-3), "unit": UNIT, "scaling": "strong"}
                if rank == 0:  # the same global batch on ONE GPU of this box: strong-scaling efficiency measured in the same run
                    try:
                        ms4_1 = quick_rate(ctx, model, diffuser, w4, w4.batch, 3, 3, dev, None, 1)
                        extra["cfg4_tcrpmhc800"].update(ms_per_step_1gpu=ms4_1, strong_scaling_efficiency=ms4_1 / (world * ms4))
                    except Exception as e:  # rank 0 must reach the barrier below whatever happens
                        extra["cfg4_tcrpmhc800"]["error_1gpu"] = repr(e)[:200]
                dist.barrier()
            if world == 8:  # cfg5: length sweep, global batch 128 over 8 GPUs = 16 per GPU
                m5, d5, c5 = build_model(True)
                for name in ("cfg5_sweep128", "cfg5_sweep256", "cfg5_sweep512", "cfg5_sweep1024"):
                    w5 = synthetic.WORKLOADS[name]
                    b0, b1 = sharding.shard_range(w5.batch, rank, world)
                    ms5 = quick_rate(c5, m5, d5, w5, b1 - b0, 5, 3, dev, dist, world)
                    extra[name] = {"global_batch": w5.batch, "batch_per_gpu": b1 - b0, "n_res": w5.n_res, "ms_per_step": ms5,
                                   "value": w5.batch * w5.n_res / (ms5 * 1e-3), "unit": UNIT, "scaling": "strong"}
                    if rank == 0:
                        try:
                            ms5_1 = quick_rate(c5, m5, d5, w5, w5.batch, 3, 3, dev, None, 1)
                            extra[name].update(ms_per_step_1gpu=ms5_1, strong_scaling_efficiency=ms5_1 / (world * ms5))
                        except Exception as e:
                            extra[name]["error_1gpu"] = repr(e)[:200]
                    dist.barrier()
        except Exception as e:  # a secondary line must never take the headline down
            extra["error"] = repr(e)[:300]

    if rank == 0:
        peaks = measured_peaks()
        hbm = peaks["hbm_gbs"]
        # ---- IPA attention, HBM-bound.  Algorithmic bytes per call (SURVEY §8d with z stored as fp16 = 256 B/pair): z read once +
        # per-residue q/k/v/points read once + concat written once.  `roofline_ipa` divides them by the time of ALL kernels that
        # implement the op (frames on points, Q.K^T, bias + softmax + o_pair, A.V, inverse frames); `roofline_ipa_core_kernel` is the
        # z-streaming kernel alone with the bytes THAT kernel must move (z + logits in + probabilities out + o_pair).
        ipa_bytes = B * (256 * N * N + 38064 * N)
        n_at, ms_at = prof["ipa_attn"]
        at_s = ms_at * 1e-3 / max(n_at, 1)
        n_ipa, ms_ipa = prof["ipa_core"]
        ipa_s = ms_ipa * 1e-3 / max(n_ipa, 1)
        core_bytes = B * (256 * N * N + 2 * 8 * 4 * N * N + 8 * 32 * 4 * N)
        n_et, ms_et = prof["edge_transition"]
        et_flops = 688128.0 * B * N * N  # 2*MACs as the reference computes them
        et_s = ms_et * 1e-3 / max(n_et, 1)
        n_fw, ms_fw = prof["forward"]
        shares = {k: (v[1] / ms_fw if ms_fw > 0 else None) for k, v in prof.items()}
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full captures (only valid for the workload they were taken on)
            for fn in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
                pth = os.path.join(ROOT, "profiles", fn)
                if os.path.exists(pth):
                    tj = json.load(open(pth))
                    if tj.get("workload") == wl.name and tj.get("batch", wl.batch) == B:
                        traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict)}
                    break
        except Exception:
            pass
        roof_ipa = {"kernel": "IPA attention: ipa_prep + Q.K^T + ipa_core + A.V + ipa_opt (all kernels of the op)", "bound": "hbm",
                    "achieved": ipa_bytes / at_s / 1e9, "peak": hbm, "unit": "GB/s", "frac": ipa_bytes / at_s / 1e9 / hbm,
                    "traffic": traffic.get("ipa_attention"), "launches": n_at, "avg_ms": at_s * 1e3, "algorithmic_bytes": ipa_bytes,
                    "share_of_forward": shares["ipa_attn"], "peak_source": peaks["source"]}
        roof_core = {"kernel": "ipa_core_kernel alone", "bound": "hbm", "achieved": core_bytes / ipa_s / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": core_bytes / ipa_s / 1e9 / hbm, "traffic": traffic.get("ipa_core_kernel"), "launches": n_ipa, "avg_ms": ipa_s * 1e3,
                     "algorithmic_bytes": core_bytes, "share_of_forward": shares["ipa_core"], "peak_source": peaks["source"]}
        tpeak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        roof_et = {"kernel": "et_fused_kernel (+ per-residue prologue GEMMs)", "bound": "tensor", "achieved": et_flops / et_s / 1e12, "peak": tpeak,
                   "unit": "TFLOP/s", "frac": et_flops / et_s / 1e12 / tpeak, "traffic": traffic.get("et_fused_kernel"), "launches": n_et, "avg_ms": et_s * 1e3,
                   "share_of_forward": shares["edge_transition"], "peak_source": peaks["source"] + " (sustained bf16)"}
        dominant = roof_et if ms_et >= ms_at else roof_ipa
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config,
                "timesteps_per_sec": K / (ms_max * 1e-3), "gpu_launches": int(launches), "clocks": clk, "diag": diag,
                "roofline": dominant, "roofline_ipa": roof_ipa, "roofline_ipa_core_kernel": roof_core, "roofline_edge_transition": roof_et,
                "time_shares_of_forward": shares}
        if e2e is not None:
            line["e2e"] = {"value": total_B * N * T_full / e2e["philox"], "unit": UNIT,
                           "h2d_bytes_per_step": int(h2d_feat / T_full), "d2h_bytes_per_step": d2h_step,
                           "what": f"framedipt_b200.inference.inference_fn, full {T_full}-step schedule incl. self-conditioning pre-pass, host feature dict in, "
                                   f"numpy [T,B,N,37,3] trajectories out (aux_traj=True), device Philox noise; wall {e2e['philox']:.3f} s",
                           "ratio_to_device_resident": (total_B * N * T_full / e2e["philox"]) / value}
            line["e2e_reference_rng_stream"] = {"value": total_B * N * T_full / e2e["numpy"], "unit": UNIT,
                                                "h2d_bytes_per_step": int(h2d_feat / T_full) + 2 * B * N * 3 * 8, "d2h_bytes_per_step": d2h_step,
                                                "what": f"same call with the reference's legacy numpy noise stream drawn on the host and uploaded; wall {e2e['numpy']:.3f} s"}
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            try:  # secondary baseline: eager PyTorch on this GPU (the reference's `use_gpu: true` path), B = 1 like the reference and the full batch
                line["eager_gpu_baseline"] = {}
                del st, pf, noise_dev
                torch.cuda.empty_cache()
                for bb in (1, B):
                    r_e, s_e = eager_gpu_rate(wl, bb, 3, 1, dev)
                    line["eager_gpu_baseline"][f"B={bb}"] = {"value": r_e, "unit": UNIT, "ms_per_step": s_e * 1e3}
                line["eager_gpu_baseline"]["what"] = ("oracle port (the reference's torch ops, eager, fp32 with TF32 off) on the same B200, host numpy "
                                                       "reverse step every timestep like experiments/utils.py:292-412; 3 timesteps after 1 warm-up")
            except Exception as e:
                line["eager_gpu_baseline"] = {"error": repr(e)[:300]}
            rate, sps = cpu_port_rate(wl, args.cpu_steps, 1, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle port, B=1 of the batch, N_res={N}, {args.cpu_steps} timesteps after 1 warm-up ({sps:.2f} s/step)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
