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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--debug-flags", type=int, default=0, help="FDPT_OPT_DEBUG_FLAGS bit set (A/B switches of include/fdpt.h)")
    ap.add_argument("--no-spinup", action="store_true", help="skip the untimed clock spin-up (for ncu launch lists, where every launch is expensive)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_tcr350")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the CPU baseline leg (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: W >= 3"

    from framedipt_b200 import synthetic

    wl = synthetic.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": f"{wl.name}: B={wl.batch}/GPU, N_res={wl.n_res}, schedule num_t={wl.num_t}, inpainting={not wl.de_novo}",
              "batch_per_gpu": wl.batch, "n_res": wl.n_res, "parallelism": f"sample-parallel x{args.gpus}",
              "l2_policy": f"inputs larger than L2: pair representation z (fp16) = {wl.batch * wl.n_res ** 2 * 256 / 1e6:.0f} MB, streamed 11x per forward"}

    if args.impl == "reference":
        if rank != 0:
            return
        rate, sps = cpu_port_rate(wl, args.steps, args.warmup, cores)
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sps * 1e3 * wl.batch, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"B=1 of the batch (the reference runs one sample at a time), N_res={wl.n_res}, {args.steps} timesteps; "
                                           f"ms_per_step is scaled to the full batch of {wl.batch}"},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    from framedipt_b200 import SE3Diffuser
    from framedipt_b200.config import default_conf
    from framedipt_b200.inference import build_schedule
    from framedipt_b200.params import synthetic_state_dict
    from framedipt_b200.score_network import ScoreNetwork

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    conf = default_conf(input_aatype=not wl.de_novo)
    diffuser = SE3Diffuser(conf.diffuser)
    model = ScoreNetwork(conf.model, diffuser, inpainting=not wl.de_novo)
    # weights: rank 0 owns them, NCCL broadcast of the packed blob over NVLink (C1 in SURVEY §7)
    sd = synthetic_state_dict(0, with_aatype=not wl.de_novo)
    if world > 1:
        from framedipt_b200 import sharding

        if rank != 0:
            sd = {k: torch.zeros_like(v) for k, v in sd.items()}
        sd = sharding.broadcast_state_dict(sd, dist, 0, dev)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    ctx = model.context(dev)

    np.random.seed(123 + rank)
    feats = synthetic.make_features(wl, diffuser, seed=0)
    B, N = wl.batch, wl.n_res
    K, W = args.steps, args.warmup
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
    if args.debug_flags:
        ctx.set_option(3, args.debug_flags)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local_rank) if rank == 0 else None  # NVML initialisation happens here, before any GPU work is timed
    # ---------------- clock spin-up (untimed, not counted as warm-up steps): a fresh box idles in a low power state and both the SM
    # and the HBM clocks ramp over the first hundreds of milliseconds of load; without this the timed pass was occasionally 30-100 %
    # slow while the later e2e pass never was.  The load is the workload itself (memory- and tensor-heavy), for >= 0.6 s. ------------
    sc, te = segment(0, W)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.6 and not args.no_spinup:
        ctx.sample(pf, sc, te, noise_dev[:W], self_condition=False)
        torch.cuda.synchronize(dev)

    # ---------------- warm-up (untimed) ----------------
    sc, te = segment(0, W)
    out = ctx.sample(pf, sc, te, noise_dev[:W], self_condition=True)
    pf.rigids_t = out["rigid_traj"][0].contiguous()
    pf.sc_ca_t = out["trans_traj"][0].contiguous()
    torch.cuda.synchronize(dev)

    # ---------------- timed region: K steps, inputs resident in HBM ----------------
    sc, te = segment(W, K)
    te = te.to(dev)
    out_buf = ctx.alloc_traj(B, N, K)  # trajectory buffers allocated before the clock starts (a cold cudaMalloc costs 1 - 60 ms)
    l0 = ctx.launch_count()
    barrier()
    if clocks:
        clocks.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cap0, wall0 = ctx.stat(0), time.perf_counter()
    ev0.record()
    wall_a = time.perf_counter()
    out = ctx.sample(pf, sc, te, noise_dev[W:], self_condition=False, out=out_buf)
    wall_b = time.perf_counter()
    ev1.record()
    wall_enqueue = time.perf_counter() - wall0
    barrier()
    ms = ev0.elapsed_time(ev1)
    # host-side sanity of the timed region: no graph capture inside it, and the enqueue (Python + C call) is a small part of it
    diag = {"graph_captures_in_timed_region": ctx.stat(0) - cap0, "enqueue_wall_ms": round(wall_enqueue * 1e3, 3),
            "sample_host_ms": {k: round(v, 3) for k, v in ctx.last_sample_host_ms.items()}}
    launches = ctx.launch_count() - l0
    clk = clocks.stop() if clocks else None
    t_all = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms_max = float(t_all.item())
    value = world * B * N * K / (ms_max * 1e-3)

    # ---------------- per-kernel timings: the same K steps once more with CUDA-event pairs around the hot kernels ----------------
    # (the event pairs sit between kernels, so this pass enqueues the step directly instead of replaying its CUDA graph)
    ctx.profile_enable(True)
    ctx.sample(pf, sc, te, noise_dev[W:], self_condition=False)
    torch.cuda.synchronize(dev)
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---------------- e2e: public API path with HOST buffers (H2D of each step's noise, D2H of its results) ----------------
    sc_e, te_e = segment(W, K)
    h_out = {"prot_traj": torch.empty(K, B, N, 5, 3).pin_memory(), "rigid_traj": torch.empty(K + 1, B, N, 7).pin_memory(),
             "trans_traj": torch.empty(K, B, N, 3).pin_memory(), "rigid_0_traj": torch.empty(K, B, N, 5, 3).pin_memory()}
    for e2e_pass in range(2):  # pass 0 is an untimed warm-up of exactly the same call (allocator blocks of these sizes, pinned staging)
        barrier()
        t0 = time.perf_counter()
        nd = noise_host[W:].to(dev, non_blocking=True)
        o2 = ctx.sample(pf, sc_e, te_e, nd, self_condition=False)
        for k in h_out:
            h_out[k].copy_(o2[k], non_blocking=True)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
    t_all = torch.tensor([e2e_s], device=dev)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        # final gather of the finished samples' coordinates (C2 in SURVEY §7), outside the step loop
        from framedipt_b200 import sharding

        gathered = sharding.gather_samples(o2["prot_traj"][0].contiguous(), dist, 0)
        assert rank != 0 or gathered.shape[0] == world * B
    e2e_value = world * B * N * K / float(t_all.item())
    h2d = int(2 * B * N * 3 * 8)
    d2h = int(B * N * (15 + 15 + 7 + 3) * 4)

    if rank == 0:
        peaks = measured_peaks()
        n_ipa, ms_ipa = prof["ipa_core"]
        # SURVEY §8d with z stored as fp16 (256 B/pair): z read once + per-residue q/k/v/points + concat (fp32)
        ipa_bytes = B * (256 * N * N + 38064 * N)
        ipa_s = ms_ipa * 1e-3 / max(n_ipa, 1)
        n_et, ms_et = prof["edge_transition"]
        et_flops = 688128.0 * B * N * N  # 2*MACs as the reference computes them
        et_s = ms_et * 1e-3 / max(n_et, 1)
        n_fw, ms_fw = prof["forward"]
        shares = {k: (v[1] / ms_fw if ms_fw > 0 else None) for k, v in prof.items()}
        dominant = "edge_transition" if ms_et >= ms_ipa else "ipa_core"
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full captures (only valid for the workload they were taken on)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
            if tj.get("workload") == wl.name:
                traffic = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if isinstance(v, dict)}
        except Exception:
            pass
        roof_ipa = {"kernel": "ipa_core_kernel", "bound": "hbm", "achieved": ipa_bytes / ipa_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ipa_bytes / ipa_s / 1e9 / peaks["hbm_gbs"], "traffic": traffic.get("ipa_core_kernel"), "launches": n_ipa, "avg_ms": ipa_s * 1e3,
                    "algorithmic_bytes": ipa_bytes,
                    "share_of_forward": shares["ipa_core"], "peak_source": peaks["source"]}
        tpeak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        roof_et = {"kernel": "et_fused_kernel (+ per-residue prologue GEMMs)", "bound": "tensor", "achieved": et_flops / et_s / 1e12, "peak": tpeak,
                   "unit": "TFLOP/s", "frac": et_flops / et_s / 1e12 / tpeak, "traffic": traffic.get("et_fused_kernel"), "launches": n_et, "avg_ms": et_s * 1e3,
                   "share_of_forward": shares["edge_transition"], "peak_source": peaks["source"] + " (sustained bf16)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "timesteps_per_sec": world * K / (ms_max * 1e-3), "gpu_launches": int(launches), "clocks": clk, "diag": diag,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "roofline": roof_et if dominant == "edge_transition" else roof_ipa,
                "roofline_ipa": roof_ipa, "roofline_edge_transition": roof_et,
                "time_shares_of_forward": shares}
        if world == 1 and not args.no_cpu_baseline:
            rate, sps = cpu_port_rate(wl, args.cpu_steps, 1, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle port, B=1 of the batch, N_res={N}, {args.cpu_steps} timesteps after 1 warm-up ({sps:.2f} s/step)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
