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

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.3)
        self.p.terminate(); self.p.wait()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
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


def sample_batch(cores):
    """Images per CPU step: the oracle parallelises over samples, so the sample must be >= the thread count to use
    all host threads, and small enough that one step stays within ~10-30 s."""
    return 128 if cores >= 32 else (64 if cores >= 16 else 16)


def cpu_baseline(B_sample, threads, steps=1, warm=0):
    """The oracle port of the reference's CPU path (no Torch7/LuaJIT exists in this image) timed on the host."""
    from oracle import pyoracle as po
    L = po.lib()
    L.og_set_threads(threads)
    rng = np.random.default_rng(1)
    G, D = po.Model(po.G32UPC, 3, 100, seed=1), po.Model(po.D32_ST3, 3, 100, seed=2)
    T = po.Trainer(G, D)
    cfg = po.default_cfg(B_sample)
    times = []
    for s in range(warm + steps):
        real = rng.uniform(0, 1, (1, B_sample // 2, 3, 32, 32)).astype(np.float32)
        zD = rng.uniform(-1, 1, (1, B_sample // 2, 100)).astype(np.float32)
        zG = rng.uniform(-1, 1, (1, B_sample, 100)).astype(np.float32)
        masks = np.stack([po.make_D_masks(B_sample, rng) for _ in range(2)])
        t0 = time.perf_counter()
        T.step(cfg, real, zD, zG, masks)
        if s >= warm:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return B_sample / dt, dt


def run_reference(args, rank, out_stream):
    if rank != 0:
        return
    from oracle import pyoracle as _po
    cores = _po.usable_cpus()
    Bs = sample_batch(cores)
    v, dt = cpu_baseline(Bs, cores, steps=max(1, min(args.steps, 3)), warm=min(args.warmup, 1))
    sample = "G32up-c+D32_st3 RGB step at batch %d (the B=128 workload cut to %d images per step), oracle port, %d OpenMP threads" % (Bs, Bs, cores)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "G32up-c + D32_st3, RGB 3x32x32, adversarial.train loop body, CPU", "batch_per_step": Bs},
           "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
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
    limit = float(os.environ.get("CATGEN_BENCH_LIMIT_S", "900"))
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
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        stage("torch.distributed init (nccl)")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            lib.check(L.cg_dist_unique_id(raw))
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        stage("broadcast ncclUniqueId")
        dist.broadcast(idbuf, 0)
        stage("cg_dist_init (ncclCommInitRank)")
        lib.check(L.cg_dist_init(rank, world, bytes(idbuf.cpu().numpy().tobytes())))

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
        lib.check(L.cg_sync())
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

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

    stage("warm-up steps")
    for i in range(W):
        dev_step(i, False)
        stage("warm-up step %d enqueued" % i)
    barrier()
    stage("timed region")
    clocks = Clocks(local) if rank == 0 else None
    L.cg_reset_launch_count()
    lib.check(L.cg_timer_start())
    t_wall = time.perf_counter()
    for i in range(W, W + K):
        dev_step(i, i == W + K - 1)      # the loss of the last step is read back: the result of the region is consumed
    ms = C.c_float()
    lib.check(L.cg_timer_stop(C.byref(ms)))
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    launches = int(L.cg_launch_count())
    clk = clocks.stop() if clocks else None
    ms_max = ms.value
    if dist is not None:
        import torch
        t = torch.tensor([ms.value], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_max = float(t.item())
    value = world * B * K / (ms_max / 1e3)
    # the same K device-resident steps again WITHOUT the nvidia-smi sampler: e2e (measured unsampled) kept coming out
    # above `value`, and the only thing specific to the region above is the 100 ms NVML polling
    barrier()
    lib.check(L.cg_timer_start())
    for i in range(W, W + K):
        dev_step(i, i == W + K - 1)
    ms_u = C.c_float(); lib.check(L.cg_timer_stop(C.byref(ms_u)))
    barrier()
    ms_u_max = ms_u.value
    if dist is not None:
        import torch
        t = torch.tensor([ms_u.value], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_u_max = float(t.item())

    stage("end-to-end region")
    # ---- end to end through the public call with HOST buffers (pinned), H2D + D2H inside the timed region
    import torch
    pin = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory()
    rng = np.random.default_rng(7 + rank)
    Ke = K
    h_real, h_zd, h_zg = pin(Ke, hB, Cc, 32, 32), pin(Ke, hB, nz), pin(Ke, B, nz)
    h_real.copy_(torch.from_numpy(rng.uniform(0, 1, h_real.shape).astype(np.float32)))
    h_zd.copy_(torch.from_numpy(rng.uniform(-1, 1, h_zd.shape).astype(np.float32)))
    h_zg.copy_(torch.from_numpy(rng.uniform(-1, 1, h_zg.shape).astype(np.float32)))
    h_out = pin(Ke, B + 2)
    outp = h_out.numpy()

    def host_step(i):
        lib.check(L.cg_train_step(T.h, C.byref(cfg), h_real[i].data_ptr(), h_zd[i].data_ptr(), h_zg[i].data_ptr(),
                                  lib.P(outp[i, B:B + 1]), lib.P(outp[i, B + 1:B + 2]), lib.P(outp[i, :B])))
    for i in range(min(3, Ke)):
        host_step(i)
    barrier()
    lib.check(L.cg_timer_start())
    for i in range(Ke):
        host_step(i)
    ms_e = C.c_float(); lib.check(L.cg_timer_stop(C.byref(ms_e)))
    barrier()
    ms_e_max = ms_e.value
    if dist is not None:
        t = torch.tensor([ms_e.value], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_e_max = float(t.item())
    e2e = world * B * Ke / (ms_e_max / 1e3)
    h2d = 4 * (n_real + n_zd + n_zg); d2h = 4 * (B + 2)
    assert np.isfinite(outp).all(), "non-finite loss / D output in the end-to-end region"

    stage("replica check / profile pass")
    # ---- data-parallel invariant: replicas start from the same seeds and apply the averaged gradient, so after any number
    # of steps every rank must hold bit-identical parameters (different data per rank, one all-reduce per network per update)
    in_sync = None
    if dist is not None:
        chk = torch.tensor([float(np.sum(G.get_params().astype(np.float64))), float(np.sum(D.get_params().astype(np.float64)))], dtype=torch.float64, device="cuda")
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))

    # ---- per-kernel durations (CUDA events around every launch) on a separate short pass: events perturb the step time
    roof = None
    kp = min(K, 3)
    # one stream, program order: with the step's concurrent lanes on, an event-bracketed launch also counts the time it
    # shares the SMs with other lanes' kernels, which is not that kernel's duration
    lib.check(L.cg_set_concurrency(0))
    if rank == 0:
        lib.check(L.cg_profile_enable(1))
    # EVERY rank runs these steps: each contains the gradient all-reduces, and a rank-0-only pass deadlocked the
    # first 2-GPU runs (rank 0 waiting in ncclAllReduce for ranks that had already finished)
    for i in range(kp):
        dev_step(W + i, False)
    barrier()
    lib.check(L.cg_set_concurrency(1))
    if rank == 0:
        buf = C.create_string_buffer(1 << 16)
        lib.check(L.cg_profile_report(buf, len(buf)))
        lib.check(L.cg_profile_enable(0))
        prof = json.loads(buf.value.decode())
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
        roof["timed"] = ("CUDA events around every launch of %d eager steps issued on ONE stream (cg_set_concurrency(0)); the timed region itself "
                         "replays the step as one CUDA graph with concurrent lanes, where a bracketed launch would also count time shared with other lanes" % kp)
        roof["top5"] = [{"kernel": k["kernel"], "share": round(k["ms"] / tot, 4), "launches_per_step": k["launches"] / kp} for k in prof[:5]]
        roof["kernels"] = [{"kernel": k["kernel"], "ms_per_step": round(k["ms"] / kp, 4), "launches_per_step": k["launches"] / kp} for k in prof[:30]]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as _po
        cores = _po.usable_cpus()
        Bs = sample_batch(cores)
        v, dt = cpu_baseline(Bs, cores, steps=1, warm=0)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "one adversarial.train loop body of the same G32up-c+D32_st3 RGB workload at batch %d (GPU arm: %d) taking %.1f s, oracle port with %d OpenMP threads; "
                         "Torch7/LuaJIT do not exist in this image" % (Bs, B, dt, cores)}

    if rank == 0:
        fl = step_flops(B)
        out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32" if L.cg_get_conv_engine() == 0 else "f16 operands / f32 accumulate on tcgen05 where the engine takes the shape, else f32",
               "data": "synthetic",
               "config": {"workload": "BASELINE configs[1] per GPU: G32up-c + D32_st3, RGB 3x32x32, batch %d per GPU (global %d), D_iterations=1, G_iterations=1, "
                                      "Adam, D_L2=1e-4, clamps 1/5, dropout on" % (B, B * world),
                          "parallelism": "dp%d" % world, "global_batch": B * world,
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
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
