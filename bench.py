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
    ap.add_argument("--no-cudnn-baseline", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the configs[2] leg (runs when N is a multiple of 4)")
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
    kind 'port'     : oracle/loops.py, the line-by-line CPU restatement (what travels to the GPU box), run with
                      weight_grads=True: FULL forward (layers after the hook, avgpool, fc included) and `cost.backward()`
                      with every backbone parameter requiring grad, exactly the work image_attacks.py:334, 351-353 does.
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
        work = "the unmodified reference class: full forward, cost.backward() incl. all weight gradients"
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
        work = ("oracle/loops.py port of the reference loop: full forward, cost.backward() with all backbone parameters "
                "requiring grad (weight gradients computed, as image_attacks.py:351-353)")
        from oracle import loops as OL
        model = backbones.seeded_random_init("resnet50", 0)
        assert all(p.requires_grad for p in model.parameters())
        hooked = [OL.HookedModel(model, "resnet", DEPTH)]
        OL.image_guided_loop(hooked, videos.numpy(), EPS, steps, STEP_SIZE, weight_grads=True,
                             tap=lambda i, d: stamps.append(time.perf_counter()))
        assert model.conv1.weight.grad is not None and model.layer2[-1].conv3.weight.grad is not None   # layers up to the hook
    dt = (stamps[-1] - stamps[0]) / (len(stamps) - 1)
    return {"value": frames / dt, "unit": UNIT, "cores": cores, "threads": torch.get_num_threads(), "kind": kind,
            "ms_per_step": dt * 1e3, "weight_grads": True,
            "sample": "1 clip x %d frames x 3x%dx%d, %d steps, I2V ResNet-50 layer2 on %d host threads; kind=%s: %s; "
                      "per-step time between first and last optimizer step" % (frames, side, side, steps, cores, kind, work)}


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
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "weight_grads", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# --------------------------------------------------------------------------------------------------
# measured TF32 tensor peak, library (cuDNN) baselines, the reference loop as-is on the GPU
# --------------------------------------------------------------------------------------------------
def measure_tf32_peak(device, n=8192, sustain_s=2.0):
    """torch.matmul on two n^3 float32 operands with allow_tf32=True (cuBLAS TF32 tensor-core GEMM), measured the way
    MEASURED_PEAKS.json measures bf16: best of 10 (burst) and back to back for `sustain_s` (sustained), 2*n^3 flops."""
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=device)
        b = torch.randn(n, n, device=device)
        c = torch.empty(n, n, device=device)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(sustain_s * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        sustained = e0.elapsed_time(e1) / reps
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    fl = 2.0 * n ** 3
    return {"burst_tflops": fl / (best * 1e-3) / 1e12, "sustained_tflops": fl / (sustained * 1e-3) / 1e12,
            "how": "torch.matmul f32 %d^3, allow_tf32=True: best of 10 / back to back for %.0f s" % (n, sustain_s)}


def _time_run(run, warm, timed):
    import torch
    for _ in range(warm):
        run.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(timed):
        run.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / timed


def cudnn_baselines(dev_videos, native_value):
    """Library baselines of the SAME step on the SAME clips, frame-steps/s on one GPU:

    truncated : comparison point (ii) — torchvision modules through cuDNN, cut at the hooked layer, input gradient only
                (i2v_b200/engine_cudnn.py), feeding the same K1 / K3a kernels; FP32 (`cudnn.allow_tf32=False`, the
                parity-equivalent arm) and TF32, NCHW and channels_last.
    as_is     : comparison point (i) — the reference's loop as image_attacks.py:294-364 runs it on a GPU: FULL forward of
                the torchvision model with a forward hook, F.cosine_similarity, `cost.backward()` with every weight
                requiring grad, torch.optim.Adam on the modifier, `print(cost)`-style host sync every step; restated here
                with torch ops only in tools/reference_on_gpu.py (no kernel of this repo on the path)."""
    import torch
    from i2v_b200 import attack_loop, backbones, engines
    out = {"unit": UNIT, "frames": int(dev_videos.shape[0] * dev_videos.shape[2]), "truncated_dgrad_only": {}, "as_is": {}}
    for name in ("cudnn", "cudnn_tf32", "cudnn_cl", "cudnn_tf32_cl"):
        eng = engines.make_engine(backbones.get_model("resnet50"), "resnet50", DEPTH, name)
        run = attack_loop.ImageGuidedRun([eng], EPS, 5, STEP_SIZE)
        run.setup(dev_videos)
        ms = _time_run(run, 2, 3)
        out["truncated_dgrad_only"][name] = {"value": run.N / (ms * 1e-3), "ms_per_step": ms, "chunk_frames": run.chunk}
        del run, eng
        torch.cuda.empty_cache()

    # ---- the reference loop as-is (4 clips = 128 frames: a full ResNet-50 autograd graph at 512 frames does not leave
    # room next to the rest of the bench; the reference's own usage is batch size 1, image_main.py:82-89) ------------
    from tools import reference_on_gpu
    vids = dev_videos[:4]
    for tag, tf32 in (("fp32", False), ("tf32", True)):
        r = reference_on_gpu.run(vids, 6, STEP_SIZE, EPS, tf32=tf32, timed_from=2)
        nfr = int(vids.shape[0] * vids.shape[2])
        out["as_is"][tag] = {"value": nfr / (r["ms_per_step"] * 1e-3), "ms_per_step": r["ms_per_step"], "frames": nfr,
                             "final_cost": float(r["cost"][-1])}
        del r
        torch.cuda.empty_cache()
    best_fp32 = max(v["value"] for k, v in out["truncated_dgrad_only"].items() if "tf32" not in k)
    best_tf32 = max(v["value"] for k, v in out["truncated_dgrad_only"].items() if "tf32" in k)
    out["native_fp32_parity_over"] = {"cudnn_truncated_fp32": native_value / best_fp32, "cudnn_truncated_tf32": native_value / best_tf32,
                                      "reference_as_is_gpu_fp32": native_value / out["as_is"]["fp32"]["value"],
                                      "reference_as_is_gpu_tf32": native_value / out["as_is"]["tf32"]["value"]}
    return out


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: ensemble, one backbone per GPU, NCCL all-reduce of the perturbation gradient
# --------------------------------------------------------------------------------------------------
ENSEMBLE_MODELS = ["resnet50", "vgg", "densenet121", "squeezenet"]


class _TimedHook:
    """dist.ReduceHook with CUDA events around the two collectives of a step."""

    def __init__(self, inner):
        import torch
        self.inner, self.torch, self.marks = inner, torch, []

    def _ev(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def grad(self, g):
        a = self._ev()
        self.inner.grad(g)
        self.marks.append([a, self._ev()])

    def cos_rows(self, cos):
        self.inner.cos_rows(cos)
        self.marks[-1].append(self._ev())


def ensemble_leg(rank, world, device, steps):
    """AENS-I2V (TPAMI_attack.py:141-320) over ENSEMBLE_MODELS, layers [2, 3] of each, placement='ensemble': rank r holds
    member r mod 4 of replica r // 4, every replica attacks its own 32-frame clip, and each step all-reduces
    dcost/dtrue_image (19.3 MB) and the [8, 32] cosine table inside the replica's NCCL group.  Reports whole-job
    frame-steps/s, the per-member compute time before the exchange (the load imbalance) and the time inside the
    collectives (for every rank but the slowest this includes waiting for it)."""
    import torch
    import torch.distributed as tdist
    import TPAMI_attack
    from i2v_b200 import attack_loop, dist as D, synth
    M = len(ENSEMBLE_MODELS)
    if world < M or world % M:
        return None
    atk = TPAMI_attack.AENS_I2V_MF(ENSEMBLE_MODELS, {n: [2, 3] for n in ENSEMBLE_MODELS}, STEP_SIZE, momentum=0.5,
                                   steps=steps, placement="ensemble")
    plan = atk._plan
    videos = synth.clip(1000 + plan.replica, b=1, f=FRAMES, h=SIDE, w=SIDE)[0].to(device)
    W = 2
    hook = _TimedHook(plan.hook())
    run = attack_loop.ImageGuidedRun(atk._engines, EPS, W + steps, STEP_SIZE, adaptive=True, coeffs=atk.coeffs, momentum=0.5,
                                     reduce_hook=hook, layer_offsets=plan.layer_offsets, n_layers_total=plan.n_layers_total)
    run.setup(videos)
    for _ in range(W):
        run.step()
    hook.marks.clear()
    D.barrier()
    torch.cuda.synchronize()
    starts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s = torch.cuda.Event(enable_timing=True); s.record(); starts.append(s)
        run.step()
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), device) / steps
    compute = sum(s.elapsed_time(m[0]) for s, m in zip(starts, hook.marks)) / steps
    coll = sum(m[0].elapsed_time(m[2]) for m in hook.marks) / steps
    mine = torch.tensor([compute, coll], device=device, dtype=torch.float64)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    tdist.all_gather(allv, mine)
    res = run.finish()
    if rank != 0:
        return None
    per = {}
    for r, t in enumerate(allv):
        per["rank%d:%s" % (r, ENSEMBLE_MODELS[r % M])] = {"compute_ms": float(t[0]), "collectives_incl_wait_ms": float(t[1])}
    comp = [float(t[0]) for t in allv]
    return {"workload": "AENS-I2V, %s, layers [2,3] each, one backbone per GPU, %d replica(s) x one %d-frame 3x%dx%d clip, "
                        "FP32 parity mode" % ("/".join(ENSEMBLE_MODELS), world // M, FRAMES, SIDE, SIDE),
            "value": (world // M) * FRAMES / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "allreduce_bytes_per_step": FRAMES * 3 * SIDE * SIDE * 4 + 2 * M * FRAMES * 4,
            "allreduce_ms_per_step": min(float(t[1]) for t in allv),
            "allreduce_share_of_step": min(float(t[1]) for t in allv) / ms,
            "member_compute_ms_max_over_min": max(comp) / max(min(comp), 1e-9), "per_rank": per,
            "final_cost": float(res.cost[-1]), "nccl_group_size": M}


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
    # The timed steps run the way the product runs them: step 0 eager, step 1 captured into a CUDA graph, the rest
    # replays (attack_loop.ImageGuidedRun.step) — no per-kernel instrumentation inside the timed region.  The per-kernel
    # CUDA-event timings behind `roofline` come from K more steps of the SAME run executed eagerly right after it
    # (events around every launch on the launching stream); their sum over the profiled steps is reported as
    # `profiled_ms_per_step` next to the graph-replayed `ms_per_step`.
    run = attack_loop.ImageGuidedRun([eng], EPS, W + 2 * K, STEP_SIZE)
    run.setup(dev_videos)
    for _ in range(W):
        run.step()
    capi.LAUNCHES.clear()
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
    launches = dict(capi.LAUNCHES)
    graph_replayed = run._graph is not None
    capi.PROFILE_EVENTS = []
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for _ in range(K):
        run.step()
    pv1.record()
    torch.cuda.synchronize()
    events, capi.PROFILE_EVENTS = capi.PROFILE_EVENTS, None
    prof_ms_local = pv0.elapsed_time(pv1)
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
                                                                        d["bytes"] / d["ms"] / 1e6, d["ms"] / prof_ms_local), file=sys.stderr)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # tensor roofline for the TF32 convolutions: the kernel runs inside a long step => sustained bf16 figure; TF32
    # issues at half the bf16 rate on the 5th-generation tensor cores (nominal 1.1 vs 2.25 PFLOP/s dense)
    tf32_meas = measure_tf32_peak(device) if rank == 0 else {"burst_tflops": 1.0, "sustained_tflops": 1.0, "how": ""}
    tf32_peak = tf32_meas["sustained_tflops"]          # the convolutions run inside a long step: sustained figure
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
             "share_of_step": k["ms"] / prof_ms_local, "traffic": traffic.get(name)}
        if k["flops"] > 0:
            tfl = k["flops"] / (k["ms"] * 1e-3) / 1e12
            mma = 3.0 if engine_name == "native" else 1.0
            r["tensor"] = {"achieved_algorithmic": tfl, "issued": tfl * mma, "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac_algorithmic": tfl / tf32_peak, "frac_issued": tfl * mma / tf32_peak,
                           "peak_source": "measured in this run: " + tf32_meas["how"] + " (sustained)",
                           "algorithmic_flops_per_launch": k["flops"] / k["launches"],
                           "note": "FP32-parity mode issues 3 TF32 MMAs per algorithmic MAC" if mma == 3.0 else "plain TF32"}
        roof_all[name] = r
    dominant = max(roof_all, key=lambda n: kern[n]["ms"]) if roof_all else None
    roofline = dict(roof_all[dominant], kernel=dominant, peak_source=peak_kind) if dominant else None
    if roofline and "tensor" in roofline and roofline["tensor"]["frac_issued"] > roofline["frac"]:
        # The convolution launches are a mix of HBM-bound (1x1) and tensor-bound (3x3, K-heavy) shapes; the roofline
        # that binds the aggregate is the one it sits closer to.  Tensor work is counted as ISSUED MMA flops — the
        # FP32-parity algorithm needs three TF32 MMAs per MAC (a_hi b_hi + a_hi b_lo + a_lo b_hi) — against the TF32
        # GEMM rate cuBLAS sustains on this GPU, measured in this run.  The HBM view stays next to it.
        t = roofline["tensor"]
        roofline = dict(roofline, bound="tensor", achieved=t["issued"], peak=t["peak"], unit="TFLOP/s", frac=t["frac_issued"],
                        peak_source=t["peak_source"],
                        hbm={"achieved": roofline["achieved"], "peak": roofline["peak"], "unit": "GB/s", "frac": roofline["frac"],
                             "peak_source": peak_kind},
                        note="issued = 3 x algorithmic flops (3xTF32 FP32-parity mode); frac_algorithmic = %.3f" % t["frac_algorithmic"])

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
        out_host.copy_(adv.contiguous(), non_blocking=True)
        del adv
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        adv = atk(host_videos, labels, names)          # H2D inside; cost log D2H inside (loss_info)
        # D2H of the result (pinned destination).  The attack returns the reference's permuted VIEW [b,c,f,h,w] of its
        # [b*f,c,h,w] frames (image_attacks.py:362-363); copying a strided view to the host takes 12.9 ms for these 308 MB,
        # one permuting copy on the device (1 ms) + a contiguous DMA 6.5 ms (tools/e2e_breakdown.py, D2H_SPLIT=1)
        out_host.copy_(adv.contiguous(), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = D.max_over_ranks(e0.elapsed_time(e1), device)
        e2e = {"value": world * N * K / (e2e_ms / 1e3), "unit": UNIT, "ms_total": e2e_ms,
               "h2d_bytes_per_step": host_videos.numel() * 4 / K, "d2h_bytes_per_step": (out_host.numel() * 4 + 4 * K) / K,
               "note": "one attack(videos, labels, names) call of K steps incl. setup, clean-feature pass, H2D of the "
                       "clips and D2H of the adversarial clips (second call: warm allocator); bytes are per call / K"}

    cudnn_base = None
    if rank == 0 and world == 1 and not args.no_cudnn_baseline:
        if "run" in locals():
            del run
        torch.cuda.empty_cache()
        cudnn_base = cudnn_baselines(dev_videos, value)
    ens = None
    if world >= len(ENSEMBLE_MODELS) and world % len(ENSEMBLE_MODELS) == 0 and not args.no_ensemble:
        if "run" in locals():
            del run
        del atk, eng
        torch.cuda.empty_cache()
        ens = ensemble_leg(rank, world, device, max(2, min(K, 10)))
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_rate(4)
        cpu_base = {k: b[k] for k in ("value", "unit", "cores", "kind", "weight_grads", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "I2V ResNet-50 layer2 cosine attack (BASELINE.json configs[1]): %d synthetic Kinetics "
                                   "clips x %d frames x 3x%dx%d per GPU, FP32 parity mode" % (clips, FRAMES, SIDE, SIDE),
                       "frames_per_gpu": N, "eps": "16/255", "step_size": STEP_SIZE, "engine": engine_name,
                       "weights": "torchvision random init, seed 0", "l2": "inputs_exceed_l2 (no flush needed)",
                       "chunk_frames": chunk_frames, "final_cost": float(res.cost[-1]) if len(res.cost) else None,
                       "timed_steps": "CUDA-graph replay of the step" if graph_replayed else "eager launches",
                       "profiled_ms_per_step": prof_ms_local / K},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum(launches.values())),
            "gpu_launches_by_kernel": launches, "roofline": roofline, "roofline_all": roof_all,
            "cpu_baseline": cpu_base, "cudnn_baseline": cudnn_base, "tf32_peak": tf32_meas, "ensemble": ens,
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
