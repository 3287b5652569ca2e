#!/usr/bin/env python
"""bench.py -- shapes/sec of LION's sampling hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One bench "step" = one full pass of the hot path over one batch: `generate_samples_vada_2prior`
= 1000 denoising steps of the global prior + 1000 of the latent-point prior (PVCNN2-AdaGN)
+ one VAE-decoder pass, for `--batch` shapes of 2048 latent points per GPU (BASELINE.json
configs[1]; at N > 1 every rank samples its own 32 shapes -> weak scaling, and the finished
point clouds are all-gathered once per pass, the path's only collective).  Weights are
key-seeded synthetic tensors (no checkpoints offline), data is synthetic noise.

Printed JSON line (rank 0): the base contract plus
  roofline      the dominant kernel (3x3x3 voxel convolution 64->64 @ 32^3, tcgen05 TF32) timed
                alone with CUDA events via lion_bench_conv; achieved = algorithmic FLOPs / time,
                peak = MEASURED_PEAKS.json bf16 burst / 2 (TF32 runs at half the bf16 rate)
  cpu_baseline  the CPU oracle (oracle/) on the host cores, bounded sample, extrapolated
  e2e           the same pass with all noise supplied from pinned HOST memory (H2D inside the
                timed region) and the point clouds copied back to pinned host memory
  gpu_baseline  (N=1, informational) a GPU port of the reference's EAGER path -- oracle/net.py on CUDA
                tensors (cuDNN/cuBLAS through torch, TF32 convs) + the reference's own pvcnn kernels from
                oracle/_ref -- on a bounded sample, run in a child process (`--impl reference-gpu`).  The
                reference's Python modules themselves cannot travel to the GPU box.
--impl reference times the reference's own CPU implementation of the path (the oracle port,
restated from the reference and pinned to its goldens) on all host threads.
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

T_STEPS = 1000
N_POINTS = 2048
GFLOP_PER_SHAPE = 59.87e3      # 1000 x 59.658 + 58.534 + 1000 x 0.154 GFLOP


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--batch", type=int, default=32, help="shapes per GPU")
    ap.add_argument("--ddpm-steps", type=int, default=T_STEPS, help="(debug) DDPM steps; the metric is defined at 1000")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_models(cfg, device):
    import torch
    from lion_b200.models.latent_points_ada_localprior import PVCNN2Prior
    from lion_b200.models.score_sde.resnet import PriorSEDrop
    from lion_b200.models.vae_adain import Model
    from tests.synth import synth_state_dict
    shp = lambda m: {k: list(v.shape) for k, v in m.state_dict().items()}
    gp = PriorSEDrop(cfg.sde, cfg.latent_pts.style_dim, cfg)
    gp.load_state_dict(synth_state_dict(shp(gp), 14))
    lp = PVCNN2Prior(cfg.sde, 1, cfg)
    lp.load_state_dict(synth_state_dict(shp(lp), 11))
    vae = Model(cfg)
    vae.decoder.load_state_dict(synth_state_dict(shp(vae.decoder), 13))
    dae = torch.nn.ModuleList([gp, lp]).to(device).eval()
    return dae, vae.to(device).eval()


# ------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_threads():
    """oneDNN / OpenMP scale poorly past a few dozen threads on B=1 work and the GPU boxes' 128
    logical cores are shared, so the CPU legs use at most 32 threads (reported as `cores`)."""
    return max(1, min(32, os.cpu_count() or 1))


def cpu_reference_sample(local_steps=1, global_steps=1):
    """Bounded sample of the reference's CPU path (oracle port, oracle/net.py): `local_steps`
    PVCNN2Prior denoising steps + `global_steps` global-prior steps at B=1.  The decoder pass
    (58.5 GFLOP, same U-Net minus the time embedding) is costed as one PVCNN2Prior step (59.7 GFLOP).
    Returns (shapes_per_sec extrapolated to 1000 + 1000 + 1 network evaluations, detail)."""
    import torch
    from oracle import net as ON
    from tests.synth import synth_state_dict
    torch.set_num_threads(cpu_threads())
    st = _CPU_STATE
    if not st:
        keys = json.load(open(os.path.join(ROOT, "tests", "golden", "keys.json")))
        st["sd_l"], st["sd_g"] = synth_state_dict(keys["prior"], 11), synth_state_dict(keys["global"], 14)
        g = torch.Generator().manual_seed(0)
        st["x"] = torch.randn(1, 8192, 1, 1, generator=g)
        st["style"] = torch.randn(1, 128, 1, 1, generator=g)
        st["t"] = torch.full((1,), 500.0)
        st["spec"] = ON.prior_spec()
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(local_steps):
            ON.prior_forward(st["sd_l"], st["spec"], st["x"], st["t"], st["style"])
        tl = (time.perf_counter() - t0) / local_steps
        t0 = time.perf_counter()
        for _ in range(global_steps):
            ON.global_prior_forward(st["sd_g"], st["style"], st["t"])
        tg = (time.perf_counter() - t0) / global_steps
    total = T_STEPS * (tl + tg) + tl           # seconds per shape (decoder ~ one more local step)
    return 1.0 / total, {"s_per_local_step": tl, "s_per_global_step": tg, "threads": cpu_threads()}


CPU_SAMPLE = ("%d PVCNN2Prior + %d global-prior denoising step(s) at B=1 on the CPU oracle (port of the reference's PyTorch "
              "path), decoder costed as one PVCNN2Prior step, extrapolated to 1000 + 1000 + 1 evaluations")


def run_gpu_reference(args):
    """--impl reference-gpu (informational, not part of the driver's contract): a GPU port of the
    reference's eager PyTorch path -- oracle/net.py on CUDA tensors (cuDNN / cuBLAS through torch, TF32
    convolutions and cudnn.benchmark as the reference runs them, utils/utils.py:472) with the reference's
    OWN point kernels (oracle/_ref/_pvcnn_backend.so).  The reference's Python modules themselves cannot
    travel to the GPU box.  Bounded sample: a few denoising steps of both priors at batch B, timed with
    CUDA events after a warm-up step, extrapolated to 1000 + 1000 + 1 network evaluations."""
    import torch
    from oracle import diffusion as OD
    from oracle import net as ON
    from oracle import ref_cuda_ops
    from tests.synth import synth_state_dict
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    ON.set_point_ops(ref_cuda_ops)
    B = args.batch
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "keys.json")))
    sd_l = {k: v.to(dev) for k, v in synth_state_dict(keys["prior"], 11).items()}
    sd_g = {k: v.to(dev) for k, v in synth_state_dict(keys["global"], 14).items()}
    spec = ON.prior_spec()
    sched = OD.make_schedule(T_STEPS, 1e-4, 0.02)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, 8192, 1, 1, device=dev, generator=g)
    xg = torch.randn(B, 128, 1, 1, device=dev, generator=g)
    style = torch.randn(B, 128, device=dev, generator=g)
    n_steps = 3

    def local_step(x, t):
        tt = torch.ones(B, device=dev) * (t + 1)
        eps = ON.prior_forward(sd_l, spec, x, tt, style)
        return OD.ddpm_step(sched, x, eps, t, torch.randn(x.shape, device=dev, generator=g))

    def global_step(xg, t):
        tt = torch.ones(B, device=dev) * (t + 1)
        eps = ON.global_prior_forward(sd_g, xg, tt)
        return OD.ddpm_step(sched, xg, eps, t, torch.randn(xg.shape, device=dev, generator=g))

    def timed(step, x0):
        with torch.no_grad():
            x1 = step(x0, T_STEPS - 1)                      # warm-up: cuDNN algorithm search, allocator
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for k in range(n_steps):
                x1 = step(x1, T_STEPS - 2 - k)
            e1.record()
            torch.cuda.synchronize(dev)
            assert torch.isfinite(x1).all()
        return max(e0.elapsed_time(e1) / 1000.0, time.perf_counter() - t0) / n_steps

    tl = timed(local_step, x)
    tg = timed(global_step, xg)
    total = T_STEPS * (tl + tg) + tl                        # seconds per batch of B shapes
    print(json.dumps({"impl": "reference-gpu", "metric": "shapes/sec (1000-step DDPM, 2048 latent pts, B=32)", "value": B / total,
                      "unit": "shapes/s", "n_gpus": 1, "kind": "port: oracle/net.py on CUDA (torch cuDNN/cuBLAS, TF32 convs, eager) + "
                      "the reference's own pvcnn kernels (oracle/_ref)",
                      "sample": "%d + %d denoising steps at batch %d after one warm-up step, extrapolated to 1000 + 1000 + 1 "
                                "network evaluations" % (n_steps, n_steps, B),
                      "detail": {"s_per_local_step": tl, "s_per_global_step": tg}}))


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(1, 1)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "shapes/sec (1000-step DDPM, 2048 latent pts, B=32)", "value": v, "unit": "shapes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * 32 / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "airplane prior, batch 32, 1000 DDPM steps, 2048 latent pts (configs[1])",
                       "timed_on": "host CPU, oracle port of the reference's PyTorch path; each step = a bounded sample"},
            "cpu_baseline": {"value": v, "unit": "shapes/s", "cores": cpu_threads(), "kind": "port", "sample": CPU_SAMPLE % (1, 1),
                             "detail": detail},
            "e2e": {"value": v, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference-gpu":
        run_gpu_reference(args)
        return
    if args.impl == "reference":
        return run_reference_arm(args)
    import torch
    import torch.distributed as dist
    from lion_b200 import _lib as L
    from lion_b200.config import default_prior_cfg
    from lion_b200.utils.diffusion_pvd import DiffusionDiscretized
    from lion_b200.trainers.train_2prior import generate_samples_vada_2prior

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path for --impl ours)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    L.lib()
    B, T = args.batch, args.ddpm_steps
    cfg = default_prior_cfg(num_steps=T)
    dae, vae = build_models(cfg, dev)
    diff = DiffusionDiscretized(cfg.sde, None, cfg)
    shape = vae.latent_shape()
    gathered = [torch.empty(B, N_POINTS, 3, device=dev) for _ in range(world)] if world > 1 else None
    launches = {"n": 0}

    def one_pass(seed):
        torch.manual_seed(seed * 1000 + rank)             # distinct noise per rank
        img, *_ = generate_samples_vada_2prior(shape, dae, diff, vae, B, False)
        launches["n"] += L.last_launches(dev)             # decoder pass (the sampling loops count themselves)
        if world > 1:
            dist.all_gather(gathered, img.contiguous())   # the single collective of the path
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing: K passes, CUDA events, max over ranks --------------------------
    for i in range(args.warmup):
        one_pass(i)
    barrier()
    launches["n"] = 0
    diff.total_gpu_launches = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        img = one_pass(100 + i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = tms.item()
    value = world * B * args.steps / (ms / 1000.0)
    n_launch = launches["n"] + diff.total_gpu_launches
    assert torch.isfinite(img).all()

    # ---- end to end: all noise from pinned host memory, result back to pinned host --------------
    e2e = None
    if not args.no_e2e:
        g = torch.Generator().manual_seed(1234 + rank)
        hn_g = torch.randn(T + 1, B, 128, 1, 1, generator=g).pin_memory()
        hn_l = torch.randn(T + 1, B, 8192, 1, 1, generator=g).pin_memory()
        hout = torch.empty(B, N_POINTS, 3).pin_memory()
        h2d = (hn_g.numel() + hn_l.numel()) * 4
        d2h = hout.numel() * 4

        # The 1.05 GB of latent-point noise is uploaded in chunks on a copy stream, last timesteps first (the
        # loop consumes z[T-1] .. z[0]); the sampling stream waits only for the chunk it is about to read, so
        # the H2D traffic overlaps the denoising steps instead of preceding them.  Everything stays inside the
        # timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        dl = torch.empty(hn_l.shape, device=dev)
        CH = 50                                                   # timesteps per chunk (13 MB)

        class StreamedNoise:
            """given_noise[1]: z[t] -> device tensor, after making the current stream wait for its upload"""

            def __init__(self):
                self.events = {}
                start = torch.cuda.Event()
                start.record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(start)
                    for lo in list(range(1, T + 1, CH))[::-1]:       # rows 1..T hold z[0..T-1]
                        hi = min(T + 1, lo + CH)
                        dl[lo:hi].copy_(hn_l[lo:hi], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        self.events[lo] = ev
                self.waited = set()

            def __getitem__(self, t):
                lo = 1 + ((t + 1 - 1) // CH) * CH
                if lo not in self.waited:
                    torch.cuda.current_stream().wait_event(self.events[lo])
                    self.waited.add(lo)
                return dl[1 + t]

        def e2e_pass():
            dg = hn_g.to(dev, non_blocking=True)
            dl[0].copy_(hn_l[0], non_blocking=True)               # x_T of the latent points
            zl = StreamedNoise()
            z_g, _ = diff.run_denoising_diffusion(dae[0], B, shape[0], given_noise=(dg[0], dg[1:]))
            z_l, _ = diff.run_denoising_diffusion(dae[1], B, shape[1], condition_input=vae.global2style(z_g),
                                                  given_noise=(dl[0], zl))
            pts = vae.sample(num_samples=B, decomposed_eps=vae.decompose_eps(vae.compose_eps([z_g, z_l])))
            if world > 1:
                dist.all_gather(gathered, pts.contiguous())
            hout.copy_(pts, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_pass()                                          # warm-up
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_pass()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * n_e2e / dt.item(), "unit": "shapes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "note": "x_T and every per-step noise tensor of both priors come from pinned host memory (the 1 GB of latent-point noise streams in 13 MB chunks on a copy stream, overlapped with the denoising steps); generated points are read back"}

    # ---- phase breakdown (one extra pass, CUDA events; diagnostic only) -------------------------
    phases = None
    if rank == 0 or world == 1:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.manual_seed(4242)
        ev[0].record()
        z_g, _ = diff.run_denoising_diffusion(dae[0], B, shape[0])
        ev[1].record()
        z_l, _ = diff.run_denoising_diffusion(dae[1], B, shape[1], condition_input=vae.global2style(z_g))
        ev[2].record()
        vae.sample(num_samples=B, decomposed_eps=vae.decompose_eps(vae.compose_eps([z_g, z_l])))
        ev[3].record()
        torch.cuda.synchronize(dev)
        phases = {"global_prior_loop_ms": ev[0].elapsed_time(ev[1]), "local_prior_loop_ms": ev[1].elapsed_time(ev[2]),
                  "decoder_ms": ev[2].elapsed_time(ev[3])}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (measured alone, CUDA events on its stream) -------------
    import ctypes as C
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst) / 2: kind::tf32 UMMA issues at half the bf16 rate" if peaks \
        else "fallback 1.59 PFLOP/s bf16 (B200_PROFILING.md) / 2"
    ms_k, fl = C.c_float(), C.c_double()
    L.check(L.lib().lion_bench_conv(L.ctx(dev), 27, 64, 64, 32, B, 20, 3, C.byref(ms_k), C.byref(fl), L.stream()), "bench_conv")
    achieved = fl.value / (ms_k.value * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "lion::tc::k_conv_tc (3x3x3 conv 64->64 @ 32^3, B=%d; 4 launches / denoising step, 49%% of FLOPs)" % B,
                "achieved": achieved, "peak": bf16_peak / 2.0, "unit": "TFLOP/s", "frac": achieved / (bf16_peak / 2.0),
                "ms_per_launch": ms_k.value, "flops_per_launch": fl.value, "peak_source": peak_src, "traffic": None}
    try:   # DRAM bytes of the same kernel from the committed ncu --set full capture (B=32 only)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_fp3_traffic.json")))
        if B == 32:
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_unit"] = "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum; algorithmic %d)" % tr["algorithmic_bytes"]
    except Exception:
        pass

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # reported on rank 0 at N=1 only
        cpu_reference_sample(1, 1)                       # warm-up (oneDNN primitive creation)
        v, detail = cpu_reference_sample(2, 2)
        cpu = {"value": v, "unit": "shapes/s", "cores": cpu_threads(), "kind": "port", "sample": CPU_SAMPLE % (2, 2), "detail": detail}

    # GPU-side reference port (informational): oracle/net.py on CUDA + the reference's own point kernels, in a
    # child process (the reference kernels exit() on a launch error; nothing there may take this line down)
    gpu_ref = None
    if world == 1 and not args.no_gpu_baseline:
        try:
            import subprocess
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu", "--batch", str(B)],
                               capture_output=True, text=True, timeout=200)
            rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
            gpu_ref = json.loads(rows[-1]) if rows else {"unavailable": (r.stderr or "no output")[-300:]}
        except Exception as e:      # noqa: BLE001
            gpu_ref = {"unavailable": repr(e)[:300]}

    line = {"metric": "shapes/sec (1000-step DDPM, 2048 latent pts, B=32)", "value": value, "unit": "shapes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (fp32 storage/accumulate)",
            "data": "synthetic",
            "config": {"workload": "airplane prior, batch %d per GPU, %d DDPM steps x (global prior + PVCNN2 latent-point prior) + VAE decoder, 2048 latent pts (BASELINE configs[1])" % (B, T),
                       "global_batch": B * world, "parallelism": "dp%d (independent shapes, one all_gather per pass)" % world,
                       "l2": "per-step working set (303 MB voxel grids) exceeds the 126 MB L2; no flush needed",
                       "weights": "key-seeded synthetic (tests/synth.py)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch,
            "ms_per_denoise_step_pair": ms / args.steps / T,
            "tensor_roofline_frac_whole_job": (value * GFLOP_PER_SHAPE / 1e3 / world / (bf16_peak / 2.0)) if T == T_STEPS else None,
            "phases": phases, "roofline": roofline, "cpu_baseline": cpu, "gpu_baseline": gpu_ref}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
