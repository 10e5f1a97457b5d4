#!/usr/bin/env python
"""bench.py -- throughput of the adversarial.train loop body (adversarial.lua:221-266) on B200.

One "step" = d D-updates + g G-updates on one batch of synthetic 3x32x32 images, exactly as SURVEY.md section 8d
defines it; images/s = B * steps/s.  Workload at N GPUs: BASELINE.json configs[1] per GPU (G32up-c + D32_st3,
RGB, batch 128), i.e. weak scaling; 8 GPUs is configs[3] (batch 1024 data parallel).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
N > 1 is launched by torchrun (one rank per GPU).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "cat-generator_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "32x32 RGB images/sec (G+D fwd+bwd)"
F_G, F_D = 2592.41e6, 373.98e6            # fwd FLOP / image
F_G_CONV = 2590.77e6                      # conv layers of G only


def step_flops(B, d=1, g=1):
    """B * [F_G*(d/2+3g) + 3*F_D*(d+g)]"""
    return B * (F_G * (d / 2 + 3 * g) + 3 * F_D * (d + g))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class Clocks:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "250"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """only samples taken from now on count (the sampler is started early so that its start-up is outside the timed region)"""
        try:
            self.skip = sum(1 for _ in open(self.f.name))
        except Exception:
            self.skip = 0

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.3)
        self.p.terminate(); self.p.wait()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines()[getattr(self, "skip", 0):]:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def conv_shape_names(B):
    """algorithmic MFLOP per launch -> layer name, for the convolutions of G32up-c + D32_st3 at batch B and B/2 (models.lua:196-228, 640-711)."""
    layers = [("G.conv1 512->512 3x3 @8", 8, 512, 512, 3), ("G.conv2 512->256 3x3 @16", 16, 512, 256, 3), ("G.conv3 256->128 5x5 @32", 32, 256, 128, 5),
              ("G.conv4 128->3 3x3 @32", 32, 128, 3, 3), ("D.trunk1 3->64 3x3 @32", 32, 3, 64, 3), ("D.trunk2 64->64 3x3 @32", 32, 64, 64, 3),
              ("D.br1-3.conv1 64->64 3x3 @16", 16, 64, 64, 3), ("D.br1-3.conv2 64->64 3x3 @8", 8, 64, 64, 3), ("D.br4.conv1 64->128 5x5 @16", 16, 64, 128, 5),
              ("D.br4.conv2 128->128 7x7 @8", 8, 128, 128, 7), ("G.linear 100->8192", 1, 100, 8192, 1), ("D.linear 20480->256", 1, 20480, 256, 1)]
    out = {}
    for n in (B, B // 2):
        for name, hw, ci, co, k in layers:
            key, label = round(2.0 * n * hw * hw * co * ci * k * k / 1e6), "%s B=%d" % (name, n)
            out[key] = out[key] + " or " + label if key in out else label      # equal work: G.conv1 at B and G.conv2 at B/2; D.trunk1 at B and G.conv4 at B/2
    return out


def _host_inputs(B, rng, d=1, g=1, C=3):
    real = rng.uniform(0, 1, (d, B // 2, C, 32, 32)).astype(np.float32)
    zD = rng.uniform(-1, 1, (d, B // 2, 100)).astype(np.float32)
    zG = rng.uniform(-1, 1, (g, B, 100)).astype(np.float32)
    return real, zD, zG


def cpu_oracle_port(B, threads, steps=1):
    """The C oracle port of the reference's Torch7 nn CPU path (per-sample im2col + hand-rolled SGEMM, OpenMP)."""
    from oracle import pyoracle as po
    L = po.lib()
    L.og_set_threads(threads)
    rng = np.random.default_rng(1)
    G, D = po.Model(po.G32UPC, 3, 100, seed=1), po.Model(po.D32_ST3, 3, 100, seed=2)
    T = po.Trainer(G, D)
    cfg = po.default_cfg(B)
    times = []
    for s in range(steps):
        real, zD, zG = _host_inputs(B, rng)
        masks = np.stack([po.make_D_masks(B, rng) for _ in range(2)])
        t0 = time.perf_counter()
        T.step(cfg, real, zD, zG, masks)
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return B / dt, dt


class _TorchCpu:
    """BASELINE.md section 3 "B-mkl": the same step with PyTorch CPU ops (oneDNN/MKL) + autograd, the stand-in for
    "Torch7 nn + optimised BLAS" (oracle/torch_step.py).  kind/C/B/d select the BASELINE config."""

    def __init__(self, threads, kind="G32UPC", C=3, B=128, d=1):
        import torch
        from oracle import pyoracle as po, torch_step as ts
        torch.set_num_threads(threads)
        self.ts, self.po, self.kind, self.C, self.B, self.d = ts, po, kind, C, B, d
        og = po.Model(po.G32UPC if kind == "G32UPC" else po.G32UP, C, 100, seed=1)
        od = po.Model(po.D32_ST3, C, 100, seed=2)
        self.G, self.D = ts.Net(np.array(og.params)), ts.Net(np.array(od.params))
        self.cfg = po.default_cfg(B, d, 1)
        self.rng = np.random.default_rng(1)

    def step(self, B=None):
        B = B or self.B
        cfg = self.po.default_cfg(B, self.d, 1)
        real, zD, zG = _host_inputs(B, self.rng, self.d, 1, self.C)
        masks = np.stack([self.po.make_D_masks(B, self.rng) for _ in range(self.d + 1)])
        t0 = time.perf_counter()
        self.ts.train_step(self.G, self.D, cfg, real, zD, zG, masks, self.kind, self.C)
        return time.perf_counter() - t0


def cpu_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip(); break
    except Exception:
        pass
    return {"nproc": os.cpu_count(), "model": model}


def run_reference(args, rank, out_stream):
    """--impl reference: the reference's CPU path on the box's host cores, same config as the GPU arm (G32up-c + D32_st3, RGB,
    batch 128 per step), EXACTLY args.steps timed steps after args.warmup untimed ones.  Torch7 cannot run here (BASELINE.md
    section 2): the timed implementation is the faster of our two CPU restatements, PyTorch-CPU (oneDNN/MKL) -- the C oracle port
    is timed for one step beside it."""
    if rank != 0:
        return
    from oracle import pyoracle as _po
    cores = _po.usable_cpus()
    B = args.batch
    T = _TorchCpu(cores, B=B)
    for _ in range(args.warmup):
        T.step()
    times = [T.step() for _ in range(args.steps)]
    dt = float(np.mean(times))
    v = B / dt
    ov, odt = cpu_oracle_port(B, cores, steps=1)
    # BASELINE configs[0] (c1): G32up grayscale, batch 16, one adversarial.train epoch of N_epoch=1000 = 125 steps; bounded sample of 10 steps
    c1 = {}
    for th in sorted({min(4, cores), cores}):
        T1 = _TorchCpu(th, kind="G32UP", C=1, B=16)
        T1.step()
        t1 = float(np.mean([T1.step() for _ in range(10)]))
        c1["threads_%d" % th] = {"s_per_step": t1, "s_per_epoch_extrapolated": t1 * 125, "ms_per_sample": 1000 * t1 * 125 / 1000, "images_per_s": 16 / t1}
    sample = ("%d timed adversarial.train loop bodies of G32up-c + D32_st3, RGB, batch %d (the GPU arm's config), PyTorch-CPU fp32 restatement "
              "(oracle/torch_step.py), %d threads" % (args.steps, B, cores))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BASELINE configs[1]: G32up-c + D32_st3, RGB 3x32x32, batch %d, D_iterations=1, G_iterations=1, CPU" % B, "global_batch": B},
           "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample, "host": cpu_info(),
                            "oracle_port": {"value": ov, "unit": "images/s", "cores": cores, "s_per_step": odt,
                                            "what": "C restatement of the Torch7 nn CPU algorithms (oracle/catgen_oracle.c), one step at batch %d" % B},
                            "c1_G32up_gray_b16_epoch": c1},
           "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out_stream.write(json.dumps(out) + "\n"); out_stream.flush()


def _quiet_stdout():
    """C libraries print to fd 1 (NCCL's version banner landed in front of the JSON line in the first 2-GPU run).  Point fd 1 at
    stderr for the lifetime of the process and hand back a writer on the real stdout for the single result line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out_stream = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch (BASELINE configs[1] = 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return run_reference(args, rank, out_stream)
    if args.warmup < 3:
        args.warmup = 3
    import faulthandler
    t_start = time.time()

    def stage(msg):
        """progress marker per rank on stderr; re-arms a stack dump so a hang names the exact line (multi-rank runs hung silently once)"""
        if world > 1 or os.environ.get("CATGEN_BENCH_TRACE"):
            sys.stderr.write("[bench rank %d +%.1fs] %s\n" % (rank, time.time() - t_start, msg)); sys.stderr.flush()
        faulthandler.cancel_dump_traceback_later()
        faulthandler.dump_traceback_later(45, exit=False)
    # hard watchdog: a deadlocked collective must end as a failed run with a message, never as a hang the driver has to kill
    limit = float(os.environ.get("CATGEN_BENCH_LIMIT_S", "300"))   # well inside the driver's per-run limit, so a hang leaves a stack
    def _watchdog():
        time.sleep(limit)
        sys.stderr.write("[bench rank %d] watchdog: not finished after %.0f s, aborting\n" % (rank, limit)); sys.stderr.flush()
        faulthandler.dump_traceback(all_threads=True)
        os._exit(3)
    threading.Thread(target=_watchdog, daemon=True).start()
    stage("loading libcatgen")
    from catgen import lib, models, adversarial
    L = lib.load()                       # raises if libcatgen.so is missing: there is no fallback path
    lib.init(local)
    dist = None
    if world > 1:
        # Control plane (rendezvous, barrier, max-over-ranks) = torch.distributed over gloo; data plane = ONE NCCL communicator, the
        # library's own (cg_dist_init), carrying the gradient all-reduces over NVLink.  Round 1 kept a second NCCL communicator
        # (torch's) in the same process and the 8-GPU run of the driver hung with rank 0 spinning in a collective.
        import torch.distributed as dist
        from catgen import dist as cgd
        stage("torch.distributed init (gloo control plane)")
        dist.init_process_group("gloo")
        raw = None
        if rank == 0:
            raw = C.create_string_buffer(128)
            lib.check(L.cg_dist_unique_id(raw))
            raw = raw.raw
        stage("broadcast ncclUniqueId")
        uid = cgd.broadcast_bytes(raw, 0, dist)
        stage("cg_dist_init (ncclCommInitRank, %d ranks)" % world)
        lib.check(L.cg_dist_init(rank, world, uid))
        if os.environ.get("CATGEN_SYNC_BN"):
            lib.check(L.cg_dist_set_sync_bn(1))

    if os.environ.get("CATGEN_PRECISION"):      # experiments: cost of the compensated forward (cg_set_precision(1)); the default run ships mode 0
        lib.check(L.cg_set_precision(int(os.environ["CATGEN_PRECISION"])))
    stage("creating models")
    B, Cc, nz = args.batch, 3, 100
    hB, img = B // 2, 3 * 1024
    G = models.create_G((Cc, 32, 32), nz, seed=1)
    D = models.create_D((Cc, 32, 32), True, seed=2)
    T = adversarial.Trainer(G, D)
    cfg = lib.default_cfg(B)
    K, W = args.steps, args.warmup
    nsteps = K + W

    def barrier():
        lib.check(L.cg_sync())          # cudaStreamSynchronize of the library's stream: every kernel and collective of this rank is done
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        return cgd.max_over_ranks(v, dist) if dist is not None else float(v)

    # ---- device-resident synthetic inputs, distinct per step and per rank (U[0,1) images, U(-1,1) noise)
    n_real, n_zd, n_zg = hB * img, hB * nz, B * nz
    real_d = L.cg_dev_alloc(4 * n_real * nsteps); zd_d = L.cg_dev_alloc(4 * n_zd * nsteps); zg_d = L.cg_dev_alloc(4 * n_zg * nsteps)
    assert real_d and zd_d and zg_d
    lib.check(L.cg_uniform_dev(real_d, n_real * nsteps, 0.0, 1.0, 1000 + rank, 0))
    lib.check(L.cg_uniform_dev(zd_d, n_zd * nsteps, -1.0, 1.0, 2000 + rank, 0))
    lib.check(L.cg_uniform_dev(zg_d, n_zg * nsteps, -1.0, 1.0, 3000 + rank, 0))
    lossD, lossG = np.zeros(1, np.float32), np.zeros(1, np.float32)

    def dev_step(i, want_loss):
        lib.check(L.cg_train_step_dev(T.h, C.byref(cfg), real_d + 4 * n_real * i, zd_d + 4 * n_zd * i, zg_d + 4 * n_zg * i,
                                      lib.P(lossD) if want_loss else None, lib.P(lossG) if want_loss else None))

    # the nvidia-smi sampler is started BEFORE the warm-up and given time to produce its first sample: its start-up (NVML init) stalled
    # launches for tens of ms when it fell inside the timed region (three of seven runs in round 2 showed 10-13 ms/step in this region
    # against 6.5-7 ms in the unsampled repeat below)
    clocks = Clocks(local) if (rank == 0 and not os.environ.get("CATGEN_BENCH_NOSMI")) else None
    if clocks is not None:
        t_w = time.time()
        while time.time() - t_w < 5.0 and os.path.getsize(clocks.f.name) == 0:
            time.sleep(0.05)
    stage("warm-up steps")
    for i in range(W):
        dev_step(i, i == 0)              # the first warm-up step also reads its loss back: that path is warm before the timed region
        stage("warm-up step %d enqueued" % i)
    barrier()
    stage("timed region")
    if clocks is not None:
        clocks.mark()
    L.cg_reset_launch_count()
    lib.check(L.cg_timer_start())
    t_wall = time.perf_counter()
    diag = os.environ.get("CATGEN_BENCH_STEPTIMES")
    t_steps = []
    for i in range(W, W + K):
        dev_step(i, i == W + K - 1)      # the loss of the last step is read back: the result of the region is consumed
        if diag:                         # diagnosis only (serialises host and device): where inside the region does the time go
            lib.check(L.cg_sync()); t_steps.append(time.perf_counter())
    if diag:
        sys.stderr.write("[bench steptimes ms] " + " ".join("%.2f" % (1e3 * (b - a)) for a, b in zip([t_wall] + t_steps[:-1], t_steps)) + "\n")
    ms = C.c_float()
    lib.check(L.cg_timer_stop(C.byref(ms)))
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    launches = int(L.cg_launch_count())
    clk = clocks.stop() if clocks else None
    ms_max = max_over_ranks(ms.value)
    value = world * B * K / (ms_max / 1e3)
    # the same K device-resident steps again WITHOUT the nvidia-smi sampler: e2e (measured unsampled) kept coming out
    # above `value`, and the only thing specific to the region above is the 100 ms NVML polling
    barrier()
    lib.check(L.cg_timer_start())
    for i in range(W, W + K):
        dev_step(i, i == W + K - 1)
    ms_u = C.c_float(); lib.check(L.cg_timer_stop(C.byref(ms_u)))
    barrier()
    ms_u_max = max_over_ranks(ms_u.value)

    stage("end-to-end region")
    # ---- end to end through the public call with HOST buffers (pinned), H2D + D2H inside the timed region
    def pin(*shape):
        """page-locked host buffer from the library (cudaMallocHost) viewed as a float32 ndarray; no torch CUDA context anywhere"""
        n = int(np.prod(shape))
        ptr = L.cg_host_alloc(4 * n)
        assert ptr, "cg_host_alloc failed"
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,)).reshape(shape), ptr
    rng = np.random.default_rng(7 + rank)
    Ke = K
    (h_real, p_real), (h_zd, p_zd), (h_zg, p_zg) = pin(Ke, hB, Cc, 32, 32), pin(Ke, hB, nz), pin(Ke, B, nz)
    h_real[...] = rng.uniform(0, 1, h_real.shape); h_zd[...] = rng.uniform(-1, 1, h_zd.shape); h_zg[...] = rng.uniform(-1, 1, h_zg.shape)
    outp, p_out = pin(Ke, B + 2)
    outp[...] = np.nan

    def host_step(i):
        lib.check(L.cg_train_step(T.h, C.byref(cfg), h_real[i].ctypes.data, h_zd[i].ctypes.data, h_zg[i].ctypes.data,
                                  lib.P(outp[i, B:B + 1]), lib.P(outp[i, B + 1:B + 2]), lib.P(outp[i, :B])))
    for i in range(min(3, Ke)):
        host_step(i)
    barrier()
    lib.check(L.cg_timer_start())
    for i in range(Ke):
        host_step(i)
    ms_e = C.c_float(); lib.check(L.cg_timer_stop(C.byref(ms_e)))
    barrier()
    ms_e_max = max_over_ranks(ms_e.value)
    e2e = world * B * Ke / (ms_e_max / 1e3)
    h2d = 4 * (n_real + n_zd + n_zg); d2h = 4 * (B + 2)
    assert np.isfinite(outp).all(), "non-finite loss / D output in the end-to-end region"

    stage("replica check / profile pass")
    # ---- data-parallel invariant: replicas start from the same seeds and apply the averaged gradient, so after any number
    # of steps every rank must hold bit-identical parameters (different data per rank, one all-reduce per network per update)
    in_sync = None
    if dist is not None:
        import torch
        chk = torch.tensor([float(np.sum(G.get_params().astype(np.float64))), float(np.sum(D.get_params().astype(np.float64)))], dtype=torch.float64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))

    # ---- per-kernel durations (CUDA events around every launch) on a separate short pass: events perturb the step time
    roof = None
    kp = min(K, 3)
    # one stream, program order: with the step's concurrent lanes on, an event-bracketed launch also counts the time it
    # shares the SMs with other lanes' kernels, which is not that kernel's duration
    # EVERY rank takes the SAME path through these steps -- profiling on (hence eager, never a graph capture) and one stream
    # everywhere: each step contains the gradient all-reduces, and in round 1 rank 0 ran them eagerly while ranks 1..7 captured a new
    # graph around the same collectives (SCALE_r01: 8-GPU run hung, GPU 0 busy, GPUs 1-7 idle).
    if os.environ.get("CATGEN_BENCH_TIMELINE") and world == 1:
        # experiments: ONE eager step with the concurrent lanes ON and an event pair around every launch -> which kernels overlap, where streams idle
        barrier()
        lib.check(L.cg_profile_enable(1))
        dev_step(W, False)
        barrier()
        tb = C.create_string_buffer(1 << 18)
        lib.check(L.cg_profile_timeline(tb, len(tb)))
        lib.check(L.cg_profile_enable(0))
        open(os.environ["CATGEN_BENCH_TIMELINE"], "w").write(tb.value.decode())
    barrier()
    lib.check(L.cg_set_concurrency(0))
    lib.check(L.cg_profile_enable(1))
    for i in range(kp):
        dev_step(W + i, False)
    barrier()
    buf = C.create_string_buffer(1 << 17)
    lib.check(L.cg_profile_report(buf, len(buf)))
    lib.check(L.cg_profile_enable(0))
    lib.check(L.cg_set_concurrency(1))
    barrier()
    if rank == 0:
        rows = json.loads(buf.value.decode())          # one row per (kernel, work per launch) = per layer shape
        prof = {}
        for r in rows:
            a = prof.setdefault(r["kernel"], {"kernel": r["kernel"], "launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            for f in ("launches", "ms", "flops", "bytes"):
                a[f] += r[f]
        prof = list(prof.values())
        tot = sum(k["ms"] for k in prof)
        prof.sort(key=lambda k: -k["ms"])
        top = prof[0]
        pk = peaks()
        conv = [k for k in prof if k["flops"] > 0]
        conv_ms, conv_fl = sum(k["ms"] for k in conv), sum(k["flops"] for k in conv)
        if top["flops"] > 0:
            ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": top["kernel"], "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": None, "peak_source": pk["source"] + " cuBLAS bf16 sustained (kernel timed inside a long step)",
                    "share_of_step": top["ms"] / tot, "launches_per_step": top["launches"] / kp,
                    "all_conv": {"achieved": conv_fl / (conv_ms * 1e-3) / 1e12, "share_of_step": conv_ms / tot}}
        else:
            ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9 if top["bytes"] else None
            roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (ach / pk["hbm_gbs"]) if ach else None,
                    "traffic": None, "peak_source": pk["source"], "share_of_step": top["ms"] / tot}
        try:   # DRAM bytes per launch of the dominant kernel's dominant shape, from the committed ncu --set full capture (not re-measured in this run)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
            roof["traffic"] = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            roof["traffic_of"] = {"layer": tj["layer"], "algorithmic_bytes": tj["algorithmic_bytes"], "source": tj["source"], "utchmma_pct_of_pipe_peak": tj["utchmma_pct_of_pipe_peak"]}
        except Exception:
            pass
        roof["timed"] = ("CUDA events around every launch of %d eager steps issued on ONE stream (cg_set_concurrency(0)); the timed region itself "
                         "replays the step as one CUDA graph with concurrent lanes, where a bracketed launch would also count time shared with other lanes" % kp)
        roof["top5"] = [{"kernel": k["kernel"], "share": round(k["ms"] / tot, 4), "launches_per_step": k["launches"] / kp} for k in prof[:5]]
        roof["kernels"] = [{"kernel": k["kernel"], "ms_per_step": round(k["ms"] / kp, 4), "launches_per_step": k["launches"] / kp} for k in prof[:30]]
        # per layer shape: the tensor-core kernels by TFLOP/s against the measured cuBLAS peak, the byte-moving kernels by GB/s against the measured
        # copy bandwidth (algorithmic bytes = what the kernel must read + write once)
        shapes = sorted(rows, key=lambda r: -r["ms"])
        names = conv_shape_names(B)
        roof["by_shape"] = [{"kernel": r["kernel"], "layer": names.get(round(r["flops"] / r["launches"] / 1e6), "?"), "us_per_launch": round(1e3 * r["ms"] / r["launches"], 1),
                             "launches_per_step": r["launches"] / kp, "tflops": round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 1),
                             "frac_of_peak": round(r["flops"] / (r["ms"] * 1e-3) / 1e12 / pk["tf_sustained"], 3)} for r in shapes if r["flops"] > 0][:12]
        roof["hbm_kernels"] = [{"kernel": r["kernel"], "us_per_launch": round(1e3 * r["ms"] / r["launches"], 1), "launches_per_step": r["launches"] / kp,
                                "gbs": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1), "frac_of_peak": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], 3)}
                               for r in shapes if r["flops"] == 0 and r["bytes"] > 0][:12]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample (~20-30 s of CPU work): 1 untimed + 2 timed steps of the SAME config (batch B) with the PyTorch-CPU restatement
        # (BASELINE.md section 3 "B-mkl"), and one step of the C oracle port at half the batch beside it
        from oracle import pyoracle as _po
        cores = _po.usable_cpus()
        Tc = _TorchCpu(cores, B=B)
        Tc.step()
        dt = float(np.mean([Tc.step() for _ in range(2)]))
        ov, odt = cpu_oracle_port(B // 2, cores, steps=1)
        cpu = {"value": B / dt, "unit": "images/s", "cores": cores, "kind": "port", "host": cpu_info(),
               "sample": "2 timed adversarial.train loop bodies of the same G32up-c+D32_st3 RGB workload at batch %d (%.1f s each), PyTorch-CPU fp32 restatement "
                         "(oracle/torch_step.py, oneDNN/MKL) with %d threads; Torch7/LuaJIT do not exist in this image" % (B, dt, cores),
               "oracle_port": {"value": ov, "unit": "images/s", "cores": cores,
                               "sample": "one step at batch %d (%.1f s) of the C restatement of the Torch7 nn CPU algorithms (oracle/catgen_oracle.c), %d OpenMP threads" % (B // 2, odt, cores)}}

    if rank == 0:
        fl = step_flops(B)
        out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32" if L.cg_get_conv_engine() == 0 else "f16 operands / f32 accumulate on tcgen05 where the engine takes the shape, else f32",
               "data": "synthetic",
               "config": {"workload": "BASELINE configs[1] per GPU: G32up-c + D32_st3, RGB 3x32x32, batch %d per GPU (global %d), D_iterations=1, G_iterations=1, "
                                      "Adam, D_L2=1e-4, clamps 1/5, dropout on" % (B, B * world),
                          "parallelism": "dp%d" % world, "global_batch": B * world, "forward_precision_mode": int(L.cg_get_precision()),
                          "l2": "no explicit flush: one step touches >1 GB of activations (>> 126 MB L2) and every step has distinct inputs",
                          "algorithmic_gflop_per_step_per_gpu": fl / 1e9},
               "achieved_tflops_per_gpu": fl * K / (ms_max / 1e3) / 1e12,
               "gpu_launches": launches, "wall_ms_per_step": wall_ms / K,
               "value_without_clock_sampler": world * B * K / (ms_u_max / 1e3),
               "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e_max / Ke},
               "replicas_in_sync": in_sync, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "last_loss": {"D": float(lossD[0]), "G": float(lossG[0])}}
        out_stream.write(json.dumps(out) + "\n"); out_stream.flush()
    stage("done")
    faulthandler.cancel_dump_traceback_later()
    L.cg_dev_free(real_d); L.cg_dev_free(zd_d); L.cg_dev_free(zg_d)
    for ptr in (p_real, p_zd, p_zg, p_out):
        L.cg_host_free(ptr)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
