#!/usr/bin/env python
"""Headline benchmark: frames/sec of the ZeroEGGS train step (60 fps, 75-joint pose frames per optimizer step).

    python bench.py --gpus N --steps K --warmup W            # N>1: launched by torchrun, one rank per GPU
    python bench.py --impl reference ...                     # the CPU implementation of the same path (oracle port)

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with the batch resident in HBM; `e2e` = the same metric
through TrainStep.step() with pinned HOST batches (H2D of every input + D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs[1] "train.py configs_v1 ... 1xB200": the reference-actual configs_v1 sizes (SURVEY.md top note)
    "train_v1": dict(B=32, T=256, H=1024, T_ex=384, desc="train.py configs_v1 (reference-actual: batch 32 x 256-frame windows, "
                     "decoder hidden 1024, example style T_ex=384 = mean of the reference's 256..512 draw)"),
    # the sizes BASELINE.json states in words for the same config
    "train_v1_stated": dict(B=16, T=120, H=512, T_ex=256, desc="train.py configs_v1 as worded in BASELINE.json (batch 16 x 120-frame windows, hidden 512)"),
}
CPU_SAMPLE = dict(B=4, T=32, T_ex=64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=0.5)
        except Exception:
            try:
                self.proc.kill()          # a lingering nvidia-smi poller would keep contending for the driver lock
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=(statistics.median(sm) if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def synth_batch(B, T, T_ex, seed, device=None, pinned=False):
    from zeggs_b200 import synth
    b = {k: torch.from_numpy(v) for k, v in synth.make_pose_windows(B, T, seed=seed).items()}
    b["audio"] = torch.from_numpy(synth.make_audio_features(B, T, seed=seed))
    b["style"] = torch.from_numpy(synth.make_style_example(B, T_ex, seed=seed))
    if pinned:
        return {k: v.pin_memory() for k, v in b.items()}
    if device is not None:
        return {k: v.to(device) for k, v in b.items()}
    return b


def build_stepper(H, device, world, seed=1234):
    from zeggs_b200 import modules, synth
    from zeggs_b200.train import TrainStep
    P = synth.make_params(H=H, seed=seed)
    ld = lambda m, pre: (m.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in P.items() if k.startswith(pre)}), m.to(device))[1]
    se = ld(modules.SpeechEncoder(81, 64, 64), "speech_encoder.")
    st = ld(modules.StyleEncoder(1134, 512, 64, type="attn", use_vae=True), "style_encoder.")
    de = ld(modules.Decoder(1134, 1131, 64, 64, H, 2), "decoder.")
    stats = synth.load_stats()
    return TrainStep(se, de, st, stats, stats["parents"], float(stats["dt"]), world_size=world), P, stats


def decoder_flops_per_frame(H, C=128):
    """MACs of one decoder step per sample (layer0, GRU l0 ih/hh, GRU l1 ih+hh, layer2)."""
    A = 1134 + C
    mac = A * H + 3 * H * (A + H) + 3 * H * H + 6 * H * H + H * 1131
    return 2 * mac


def run_ours(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lrank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    device = torch.device("cuda", lrank)
    torch.cuda.set_device(device)
    try:
        os.nice(-10)                      # the launch thread competes with other tenants' all-core CPU jobs on a shared node
    except Exception:
        pass
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        torch.distributed.barrier()
    from zeggs_b200 import _lib, ops
    lib = _lib.lib()
    ops.set_decoder_engine(args.engine)
    wl = WORKLOADS[args.workload]
    B, T, H, T_ex = wl["B"], wl["T"], wl["H"], wl["T_ex"]
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def measure(B, T, H, T_ex, K, W, timing=False):
        stepper, P, stats = build_stepper(H, device, world)
        dbatch = synth_batch(B, T, T_ex, seed=100 + rank, device=device)
        for _ in range(W):
            stepper.step(dbatch)
        barrier()
        if timing:
            lib.zeggs_timing_reset(); lib.zeggs_timing_enable(1)
        sampler = ClockSampler(lrank)
        if rank == 0 and timing:
            sampler.start()
        if timing:
            # nvidia-smi needs ~0.1-0.3 s to start reporting: every rank keeps its GPU under the same load meanwhile (untimed)
            for _ in range(24):
                stepper.step(dbatch)
            barrier()
            lib.zeggs_timing_reset()
        n0 = lib.zeggs_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            loss = stepper.step(dbatch)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = int(lib.zeggs_launch_count() - n0)
        clocks = sampler.stop() if (rank == 0 and timing) else None
        if clocks is not None:
            clocks["note"] = "nvidia-smi every 200 ms from the start of 24 untimed warm steps of the same loop through the timed region"
        spans = {}
        if timing:
            lib.zeggs_timing_enable(0)
            import ctypes as C
            for name in ("decoder_fwd", "decoder_bwd", "decoder_wgrad", "loss", "encoders_fwd", "encoders_bwd"):
                tot, cnt = C.c_double(0), C.c_int(0)
                lib.zeggs_timing_read(name.encode(), C.byref(tot), C.byref(cnt))
                spans[name] = dict(ms_per_step=tot.value / max(K, 1), launches=cnt.value)
        # end to end: pinned host batch -> H2D every step, loss read back every step
        hbatch = synth_batch(B, T, T_ex, seed=200 + rank, pinned=True)
        h2d = sum(v.numel() * v.element_size() for v in hbatch.values())
        # the public input pipeline (zeggs_b200.data.DevicePrefetcher, used by train()): every step's batch is copied from pinned
        # host memory inside the timed region, on a side stream, while the previous step's kernels run
        from zeggs_b200.data import DevicePrefetcher
        pf = DevicePrefetcher(device)
        for _ in range(2):
            stepper.step(pf.acquire(pf.upload(hbatch))).item()
        barrier()
        t0 = time.perf_counter()
        tok = pf.upload(hbatch)
        for i in range(K):
            db = pf.acquire(tok)
            if i + 1 < K:
                tok = pf.upload(hbatch)
            float(stepper.step(db).item())
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        barrier()
        return dict(ms=ms, e2e_ms=e2e_ms, launches=launches, clocks=clocks, spans=spans, h2d=h2d, loss=float(loss.item()), P=P, stats=stats)

    r = measure(B, T, H, T_ex, K, W, timing=True)
    frames = world * B * T
    value = frames / (r["ms"] / K) * 1e3
    peaks = load_peaks()
    # roofline of the dominant kernel (the slower of the two persistent recurrence kernels): algorithmic FLOPs = the
    # (transposed) GEMMs of every step
    fl_step = decoder_flops_per_frame(H) - 2 * 128 * 4 * H          # the speech/style columns are hoisted out of the recurrence
    dom = "decoder_bwd" if r["spans"]["decoder_bwd"]["ms_per_step"] >= r["spans"]["decoder_fwd"]["ms_per_step"] else "decoder_fwd"
    dom_ms = r["spans"][dom]["ms_per_step"]
    achieved = fl_step * B * (T - 1) / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    wbytes = (fl_step / 2) * 4                                      # fp32 weights streamed per step (L2 resident)
    roofline = dict(bound="tensor", kernel=(f"{dom}_tc_kernel" if args.engine == "tc" else f"{dom}_kernel"), achieved=round(achieved, 3), peak=peaks["bf16_tflops_sustained"],
                    unit="TFLOP/s", frac=round(achieved / peaks["bf16_tflops_sustained"], 5), traffic=None,
                    peak_source=peaks["src"] + " (cuBLAS bf16, sustained)", ms_per_launch=round(dom_ms, 3),
                    weight_stream_gbs=round(wbytes * (T - 1) / (dom_ms * 1e-3) / 1e9, 1),
                    note=("decoder forward and BPTT recurrences on tcgen05 (bf16 operands from smem images, fp32 accumulators in TMEM, fp32 state; "
                          "forward: layer 2 folded into the next step's input GEMM -> 3 all-to-all exchanges per step, M=64 MMAs, one mbarrier wait per 16 MMAs); "
                          "weight-gradient GEMMs tcgen05 bf16, encoder GEMMs tcgen05 split-bf16; B=32 rows per step: the recurrences are bound by operand "
                          "delivery per SM, grid-barrier latency and MMA issue, not by tensor peak; "
                          if args.engine == "tc" else "fp32 SIMT recurrence; batched GEMMs tcgen05 split-bf16; ") +
                         "per-step arithmetic intensity at B=32 is 16-32 FLOP/B (weight streaming from L2), see DESIGN.md")
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp) and args.workload == "train_v1":
        try:
            tr = json.load(open(tp)).get(roofline["kernel"])
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
        except Exception:
            pass
    out = dict(metric="frames/sec (train step, 60fps 75-joint pose)", value=round(value, 1), unit="frames/s", n_gpus=world, steps=K, warmup=W,
               ms_per_step=round(r["ms"] / K, 3), higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype=("bf16" if args.engine == "tc" else "f32"), data="synthetic",
               config=dict(workload=wl["desc"], per_gpu_batch=B, global_batch=world * B, window=T, hidden=H, style_example_len=T_ex,
                           parallelism=f"dp{world}", decoder_engine=args.engine, l2="per-step working set ~1.4 GB of saved activations >> 126 MB L2 (no explicit flush)"),
               e2e=dict(value=round(frames / (r["e2e_ms"] / K) * 1e3, 1), unit="frames/s", h2d_bytes_per_step=r["h2d"], d2h_bytes_per_step=4),
               gpu_launches=r["launches"], clocks=r["clocks"], roofline=roofline,
               kernel_ms_per_step={k: round(v["ms_per_step"], 3) for k, v in r["spans"].items()}, loss=r["loss"])
    if args.alt and args.workload == "train_v1":
        a = WORKLOADS["train_v1_stated"]
        ra = measure(a["B"], a["T"], a["H"], a["T_ex"], K, W)
        out["alt_config"] = dict(workload=a["desc"], value=round(world * a["B"] * a["T"] / (ra["ms"] / K) * 1e3, 1), unit="frames/s",
                                 ms_per_step=round(ra["ms"] / K, 3),
                                 e2e=round(world * a["B"] * a["T"] / (ra["e2e_ms"] / K) * 1e3, 1))
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.train_oracle import time_cpu_train_step
        cb = synth_batch(CPU_SAMPLE["B"], CPU_SAMPLE["T"], CPU_SAMPLE["T_ex"], seed=300)
        fps, n, th = time_cpu_train_step(r["P"], r["stats"], cb, budget_s=20.0)
        out["cpu_baseline"] = dict(value=round(fps, 1), unit="frames/s", cores=th, kind="port",
                                   sample=f"{n} oracle train steps at B={CPU_SAMPLE['B']}, T={CPU_SAMPLE['T']}, H={H}, T_ex={CPU_SAMPLE['T_ex']} "
                                          f"(frames/s is ~linear in B*T; os.cpu_count()={os.cpu_count()})")
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_reference(args):
    """The CPU implementation of the same path (oracle port; the Python reference cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.train_oracle import OracleTrainer
    from zeggs_b200 import synth
    wl = WORKLOADS[args.workload]
    H = wl["H"]
    torch.set_num_threads(os.cpu_count() or 1)
    P = synth.make_params(H=H, seed=1234)
    stats = synth.load_stats()
    tr = OracleTrainer(P, stats)
    b = synth_batch(CPU_SAMPLE["B"], CPU_SAMPLE["T"], CPU_SAMPLE["T_ex"], seed=300)
    K, W = args.steps, args.warmup
    for _ in range(W):
        tr.step(b)
    t0 = time.perf_counter()
    for _ in range(K):
        tr.step(b)
    dt = time.perf_counter() - t0
    fps = CPU_SAMPLE["B"] * CPU_SAMPLE["T"] * K / dt
    sample = f"each step = one CPU train step at B={CPU_SAMPLE['B']}, T={CPU_SAMPLE['T']}, H={H}, T_ex={CPU_SAMPLE['T_ex']} (bounded sample of the workload)"
    print(json.dumps(dict(impl="reference", metric="frames/sec (train step, 60fps 75-joint pose)", value=round(fps, 1), unit="frames/s",
                          n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=round(dt / K * 1e3, 2), higher_is_better=True, scaling="weak",
                          vs_baseline=None, dtype="f32", data="synthetic",
                          config=dict(workload=wl["desc"], sample=sample),
                          cpu_baseline=dict(value=round(fps, 1), unit="frames/s", cores=torch.get_num_threads(), kind="port", sample=sample),
                          e2e=dict(value=round(fps, 1), unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train_v1", choices=list(WORKLOADS))
    ap.add_argument("--alt", type=int, default=1, help="also time the BASELINE.json-worded sizes (reported under alt_config)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--engine", default="tc", choices=["tc", "fp32"], help="decoder recurrence engine: tcgen05 bf16 (default) or fp32 SIMT")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
