#!/usr/bin/env python
"""bench.py — attack frame-steps/s of the I2V per-step attack loop on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--engine cudnn|native|...] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): I2V single-layer cosine attack, ResNet-50 layer2 (random init),
16 synthetic Kinetics clips = 512 frames of 3x224x224 per GPU, FP32-parity mode, eps = 16/255,
step_size = 0.005.  A "step" is one attack step (compose -> forward -> cosine loss/grad -> backward
-> Adam) over the whole batch; the metric counts frames x steps per second.  Multi-GPU runs give
every rank its own batch (clips are independent units: no data-path collective) => weak scaling.

Timing: W warm-up steps, then exactly K steps between barrier + cuda synchronize, CUDA events on the
launching stream, MAX over ranks.  The working set of one step (512 frames x ~125 MB of activations)
exceeds the 126 MB L2 many times over, so no L2 flush is needed between iterations.

One JSON line on stdout (rank 0).  Extra objects:
  roofline     — the dominant kernel of THIS repo inside the timed region, timed live with CUDA events
  cpu_baseline — the reference's arithmetic on the host cores (oracle port; the reference classes
                 themselves when /root/reference is present), bounded sample, rank 0 at N=1 only
  e2e          — the same metric through the public drop-in API with HOST (pinned) input clips:
                 H2D of the clips and D2H of the adversarial clips + cost log inside the timed region
"""
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
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line (NCCL's version banner goes to stderr)

METRIC = "attack_frame_steps_per_sec"
UNIT = "frame-steps/s"
CLIPS_PER_GPU = 16
FRAMES, SIDE = 32, 224
STEP_SIZE, EPS = 0.005, 16 / 255
MODEL, DEPTH = "resnet50", 2
FEAT_D = 512 * 28 * 28


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default=None, help="native | native_tf32 | cudnn | cudnn_tf32 (default: $I2V_ENGINE or native)")
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU (default 16 = 512 frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shapes", action="store_true", help="print a per-shape timing table of the convolutions to stderr")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, frames=32, side=SIDE, budget_s=40.0):
    """frame-steps/s of the reference's I2V loop on the host cores, all threads.

    kind 'reference': the unmodified class from /root/reference (build container only);
    kind 'port'     : oracle/loops.py, the line-by-line CPU restatement (what travels to the GPU box).
    Both run the reference's FULL forward + autograd with weight gradients, as the reference does.
    The sample is one config-1 clip (32 frames of 3x224x224); the per-step time is measured between the
    first and the last optimizer step so that model construction and the clean-feature pass are excluded.
    """
    import torch
    from i2v_b200 import backbones, synth
    from oracle import load_reference as LR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    videos, labels = synth.clip(0, b=1, f=frames, h=side, w=side)
    stamps = []
    steps = max(2, steps)
    if LR.available():
        kind = "reference"
        ref = LR.load()
        orig = torch.optim.Adam

        class Stamped(orig):
            def step(self, closure=None):
                out = super().step(closure)
                stamps.append(time.perf_counter())
                return out
        torch.optim.Adam = Stamped
        try:
            with LR.quiet():
                atk = ref.image_attacks.ImageGuidedFMDirection_Adam(["resnet"], depth=DEPTH, step_size=STEP_SIZE, steps=steps)
                atk(videos, labels, ["clip0"])
        finally:
            torch.optim.Adam = orig
    else:
        kind = "port"
        from oracle import loops as OL
        model = backbones.seeded_random_init("resnet50", 0)
        hooked = [OL.HookedModel(model, "resnet", DEPTH)]
        OL.image_guided_loop(hooked, videos.numpy(), EPS, steps, STEP_SIZE,
                             tap=lambda i, d: stamps.append(time.perf_counter()))
    dt = (stamps[-1] - stamps[0]) / (len(stamps) - 1)
    return {"value": frames / dt, "unit": UNIT, "cores": cores, "threads": torch.get_num_threads(), "kind": kind,
            "ms_per_step": dt * 1e3,
            "sample": "1 clip x %d frames x 3x%dx%d, %d steps, I2V ResNet-50 layer2 (full forward + weight grads as the "
                      "reference runs it), per-step time between first and last optimizer step" % (frames, side, side, steps)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(2, min(args.steps, 8))
    base = cpu_reference_rate(steps + 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "I2V ResNet-50 layer2 cosine attack, CPU reference path, bounded sample: " + base["sample"],
                   "eps": "16/255", "step_size": STEP_SIZE},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    from i2v_b200 import attack_loop, backbones, capi, dist as D, engines, synth
    import image_attacks

    rank, local_rank, world = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: i2v_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    capi.device_check(device)
    backbones.set_weight_policy("random", 0)
    engine_name = engines.resolve(args.engine)

    clips = args.clips
    K, W = args.steps, args.warmup
    N = clips * FRAMES
    # every rank gets its own clips: global clip index = rank * clips + i
    videos = torch.cat([synth.clip(rank * clips + i, b=1, f=FRAMES, h=SIDE, w=SIDE)[0] for i in range(clips)], 0)
    host_videos = videos.pin_memory()
    dev_videos = host_videos.to(device)

    # the public drop-in constructor builds the backbone and its engine; both arms below share it
    atk = image_attacks.ImageGuidedFMDirection_Adam([MODEL], depth=DEPTH, step_size=STEP_SIZE, epsilon=EPS, steps=K,
                                                    engine=engine_name)
    eng = atk._engine

    # ---- device-resident arm: W warm-up steps + exactly K timed steps of one attack call --------------
    run = attack_loop.ImageGuidedRun([eng], EPS, W + K, STEP_SIZE)
    run.setup(dev_videos)
    for _ in range(W):
        run.step()
    capi.LAUNCHES.clear()
    capi.PROFILE_EVENTS = []
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    D.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        run.step()
    ev1.record()
    torch.cuda.synchronize()
    D.barrier()
    clocks = sampler.stop() if sampler else None
    ms_local = ev0.elapsed_time(ev1)
    events, capi.PROFILE_EVENTS = capi.PROFILE_EVENTS, None
    launches = dict(capi.LAUNCHES)
    res = run.finish()
    chunk_frames = run.chunk
    ms = D.max_over_ranks(ms_local, device)
    value = world * N * K / (ms / 1e3)

    # ---- per-kernel device time of OUR kernels inside the timed region (CUDA events, same stream) -----
    kern, by_shape = {}, {}
    for name, e0, e1, nbytes, flops, detail in events:
        k = kern.setdefault(name, {"ms": 0.0, "launches": 0, "bytes": 0, "flops": 0.0})
        dt = e0.elapsed_time(e1)
        k["ms"] += dt
        k["launches"] += 1
        k["bytes"] += nbytes
        k["flops"] += flops
        if detail is not None:
            d = by_shape.setdefault(detail, {"ms": 0.0, "launches": 0, "bytes": 0})
            d["ms"] += dt; d["launches"] += 1; d["bytes"] += nbytes
    if args.shapes and rank == 0:       # per-shape table of the tensor-core convolutions, in-situ (stderr)
        for detail, d in sorted(by_shape.items(), key=lambda kv: -kv[1]["ms"]):
            print("%-44s n=%-5d avg %7.1f us  %6.0f GB/s  share %.3f" % (detail, d["launches"], 1e3 * d["ms"] / d["launches"],
                                                                        d["bytes"] / d["ms"] / 1e6, d["ms"] / ms_local), file=sys.stderr)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # tensor roofline for the TF32 convolutions: the kernel runs inside a long step => sustained bf16 figure; TF32
    # issues at half the bf16 rate on the 5th-generation tensor cores (nominal 1.1 vs 2.25 PFLOP/s dense)
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    tf32_peak = bf16_peak / 2
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        pass
    roof_all = {}
    for name, k in kern.items():
        if k["ms"] <= 0:
            continue
        gbs = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
             "launches": k["launches"], "avg_us": 1e3 * k["ms"] / k["launches"],
             "algorithmic_bytes_per_launch": k["bytes"] / k["launches"],
             "share_of_step": k["ms"] / ms_local, "traffic": traffic.get(name)}
        if k["flops"] > 0:
            tfl = k["flops"] / (k["ms"] * 1e-3) / 1e12
            mma = 3.0 if engine_name == "native" else 1.0
            r["tensor"] = {"achieved_algorithmic": tfl, "issued": tfl * mma, "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac_algorithmic": tfl / tf32_peak, "frac_issued": tfl * mma / tf32_peak,
                           "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 issue rate)",
                           "algorithmic_flops_per_launch": k["flops"] / k["launches"],
                           "note": "FP32-parity mode issues 3 TF32 MMAs per algorithmic MAC" if mma == 3.0 else "plain TF32"}
        roof_all[name] = r
    dominant = max(roof_all, key=lambda n: kern[n]["ms"]) if roof_all else None
    roofline = dict(roof_all[dominant], kernel=dominant, peak_source=peak_kind) if dominant else None

    # ---- e2e arm: public drop-in API, host clips in, adversarial clips + cost log out -----------------
    e2e = None
    if not args.no_e2e:
        names = ["clip%d" % i for i in range(clips)]
        labels = torch.zeros(clips, dtype=torch.long)
        out_host = torch.empty(videos.shape, dtype=torch.float32).pin_memory()
        del run                                        # the device-resident arm's state goes back to the allocator
        # one untimed call first: a sweep attacks clip batch after clip batch, so the steady state has a warm caching
        # allocator (the first call pays ~0.3 s of cudaMalloc for 4 GB of per-call state)
        adv = atk(host_videos, labels, names)
        out_host.copy_(adv, non_blocking=True)
        del adv
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        adv = atk(host_videos, labels, names)          # H2D inside; cost log D2H inside (loss_info)
        out_host.copy_(adv, non_blocking=True)         # D2H of the result (pinned destination)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = D.max_over_ranks(e0.elapsed_time(e1), device)
        e2e = {"value": world * N * K / (e2e_ms / 1e3), "unit": UNIT, "ms_total": e2e_ms,
               "h2d_bytes_per_step": host_videos.numel() * 4 / K, "d2h_bytes_per_step": (out_host.numel() * 4 + 4 * K) / K,
               "note": "one attack(videos, labels, names) call of K steps incl. setup, clean-feature pass, H2D of the "
                       "clips and D2H of the adversarial clips (second call: warm allocator); bytes are per call / K"}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_rate(4)
        cpu_base = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "I2V ResNet-50 layer2 cosine attack (BASELINE.json configs[1]): %d synthetic Kinetics "
                                   "clips x %d frames x 3x%dx%d per GPU, FP32 parity mode" % (clips, FRAMES, SIDE, SIDE),
                       "frames_per_gpu": N, "eps": "16/255", "step_size": STEP_SIZE, "engine": engine_name,
                       "weights": "torchvision random init, seed 0", "l2": "inputs_exceed_l2 (no flush needed)",
                       "chunk_frames": chunk_frames, "final_cost": float(res.cost[-1]) if len(res.cost) else None},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum(launches.values())),
            "gpu_launches_by_kernel": launches, "roofline": roofline, "roofline_all": roof_all,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
