#!/usr/bin/env python
"""bench.py -- corrupted-image evals/sec of the corruption-sweep hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): ResNet-18 random-init, 32x32 CIFAR-shape synthetic images,
MC-dropout T=20, corruption x severity sweep with ECE / entropy / AUROC aggregates.  One *step* = one
block of `block` images taken through one (corruption, severity) cell: corrupt+normalize -> ResNet-18 x T ->
uncertainty epilogue -> histogram accumulation; consecutive steps walk the cell grid and the image blocks.
One *eval* = one (image, corruption, severity) triple through all T passes.

value      : evals/s, whole job over all ranks, inputs resident in HBM when the timed region starts
e2e        : evals/s through the public streaming API (CorruptionSweep.run_stream on HOST uint8 blocks): every step's
             pinned H2D copy of its block (prefetched on a side stream while the previous block computes) and the D2H
             read of the step's histogram arena row are inside the timed region
roofline   : the tensor-core conv kernel; achieved = algorithmic FLOPs (2*(2.408+T*34.60) MFLOP
             per eval) / summed conv-kernel device time per step (CUDA events around every conv launch)
cpu_baseline: the oracle (plain PyTorch fp32 restatement; the reference ships no code for this path) timed on
             the box's host cores on a bounded sample of the same workload
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

T_PASSES = 20
P_DROP = 0.2
TAU = 0.9
BLOCK = 4096           # images per step; measured on B200 (short runs): 512 -> 0.71 M, 1024 -> 0.85 M, 2048 -> 0.92 M, 4096 -> 0.94 M evals/s
                       # (a step has ~0.2 ms of per-kernel ramp-up/drain that a larger block amortises), profiles/r01m_block_sweep.txt
N_IMAGES = 16384
MFLOP_PREFIX, MFLOP_PASS = 2.408448, 34.608128          # MMAC per image (oracle.model.count_macs); x2 for FLOPs


def flops_per_eval(T):
    return 2.0e6 * (MFLOP_PREFIX + T * MFLOP_PASS)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.rows, self.proc = device_index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        inside = [r for (ts, r) in self.rows if self.t0 is None or (self.t0 <= ts <= (self.t1 or ts) + 0.05)]
        if len(inside) < 2:          # very short timed region: fall back to everything sampled while the GPU was busy
            inside = [r for (_, r) in self.rows]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
                try:
                    pw.append(float(f[3]))
                except ValueError:
                    pass
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


_REAL_STDOUT = None


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def cpu_baseline(sample_images=128, T=T_PASSES, repeats=1):
    """Oracle (kind 'port': the reference has no implementation of this path) on the host cores."""
    import numpy as np
    import torch
    from oracle import model as OM, philox as px, sweep as OS
    folded = OM.fold_resnet(OM.build_torchvision("resnet18", 10, 0, logit_gain=8.0))
    x = px.synthetic_images(sample_images, 32, 32, 0)
    y = px.synthetic_labels(sample_images, 10, 0)
    best, best_thr = None, None
    cand = sorted({1, max(1, (os.cpu_count() or 1) // 2), os.cpu_count() or 1})
    for thr in cand:
        torch.set_num_threads(thr)
        OS.eval_cell(folded, x[:8], y[:8], "gaussian_noise", 3, T=2)           # warm-up
        t0 = time.perf_counter()
        for _ in range(repeats):
            OS.eval_cell(folded, x, y, "gaussian_noise", 3, T=T, p=P_DROP, tau=TAU)
        dt = (time.perf_counter() - t0) / repeats
        if best is None or dt < best:
            best, best_thr = dt, thr
    try:
        load = os.getloadavg()[0]
    except OSError:
        load = None
    return {"value": sample_images / best, "unit": "evals/s", "cores": best_thr, "kind": "port",
            "sample": f"{sample_images} images x 1 cell (gaussian_noise s3) x T={T}, fp32 PyTorch oracle, best of "
                      f"threads {cand}; os.cpu_count()={os.cpu_count()}, loadavg={load}"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import model as OM, philox as px, sweep as OS
    per_step = 64            # bounded sample of a step (the GPU arm's step is BLOCK images): ~0.3 s of CPU work per step
    folded = OM.fold_resnet(OM.build_torchvision("resnet18", 10, 0, logit_gain=8.0))
    torch.set_num_threads(os.cpu_count() or 1)
    x = px.synthetic_images(per_step * 4, 32, 32, 0)
    y = px.synthetic_labels(per_step * 4, 10, 0)
    cells = [("gaussian_noise", 3), ("contrast", 2), ("impulse_noise", 4), ("brightness", 5)]

    def step(i):
        lo = (i % 4) * per_step
        c, s = cells[i % len(cells)]
        OS.eval_cell(folded, x[lo:lo + per_step], y[lo:lo + per_step], c, s, T=T_PASSES, p=P_DROP, tau=TAU,
                     first_image=lo)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    sample = (f"{per_step} of the {BLOCK} images of a step x T={T_PASSES}, cells {cells} in rotation, fp32 PyTorch oracle, "
              f"{torch.get_num_threads()} threads")
    emit_json({
        "impl": "reference", "metric": "corrupted-image evals/sec", "value": v, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, BLOCK), "sample_images_per_step": per_step,
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def workload_config(n_gpus, block):
    import fav
    return {"workload": "C2: ResNet-18 random-init (torchvision, logit-gain fixture 8.0), 32x32 CIFAR-shape, "
                        f"MC-dropout T={T_PASSES} p={P_DROP}, sweep over all {len(fav.IMPLEMENTED)} corruptions x 5 severities "
                        f"({len(fav.IMPLEMENTED) * 5} cells), ECE(15 bins)+entropy+MI+AUROC(4096 buckets)",
            "images_per_step": block, "passes": T_PASSES, "num_classes": 10,
            "corruptions": list(fav.IMPLEMENTED), "parallelism": f"image-block x cell sharding over {n_gpus} GPU(s)",
            "l2": "working set (5 activation buffers x 672 MB per step) exceeds the 126 MB L2; image blocks rotate"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: library banners (e.g. "NCCL version ...") are sent to stderr instead
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import fav
    from fav import _lib
    from fav.sweep import CorruptionSweep, SweepConfig
    import ctypes as C

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback (use --impl reference)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = args.steps

    cfg = SweepConfig(T=T_PASSES, p_drop=P_DROP, tau=TAU, logit_gain=8.0, block=BLOCK, seed=0)
    sweep = CorruptionSweep(cfg, device=local)
    sweep.prepare()
    clf = sweep.clf
    lib, h = clf.lib, clf.handle.h
    # synthetic inputs generated on the device from the Philox "images"/"labels" streams; each rank owns a
    # disjoint range of global image indices (weak scaling: fixed work per GPU)
    first = rank * N_IMAGES
    images = torch.empty((N_IMAGES, 32, 32, 3), dtype=torch.uint8, device=clf.device)
    labels = torch.empty(N_IMAGES, dtype=torch.int32, device=clf.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.fav_synth_images(h, C.c_void_p(images.data_ptr()), N_IMAGES, 32, 32, 0, first, st), "synth")
    _lib.check(lib.fav_synth_labels(h, C.c_void_p(labels.data_ptr()), N_IMAGES, 10, 0, first, st), "synth")
    items = sweep.work_items(N_IMAGES)              # (cell, block) round robin over the grid
    host_images = images.cpu().pin_memory()
    host_labels = labels.cpu().pin_memory()

    def step_resident(i):
        return sweep.run_item(images, labels, items[i % len(items)], first)

    host_rows = []

    def run_e2e(steps):
        """`steps` steps through the public streaming API on HOST blocks: per step one pinned H2D copy of the block
        (prefetched on a side stream while the previous block computes) and one D2H read of the cell's arena row."""
        seq = [items[i % len(items)] for i in range(steps)]
        host_rows.clear()
        return sweep.run_stream(host_images, host_labels, seq, first, on_row=lambda item, row: host_rows.append(int(row[0])))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        evals = 0
        if fn is run_e2e:
            evals = run_e2e(steps)
        else:
            for i in range(steps):
                evals += fn(i)
        if world > 1:
            sweep.acc.allreduce()                           # the path's one exchange, inside the timed region
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device=clf.device)
        n = torch.tensor([evals], dtype=torch.int64, device=clf.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
        return float(t.item()), int(n.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W):
        step_resident(i)
    run_e2e(W)
    sweep.reset()

    l0 = clf.handle.launches()
    sampler.mark_begin()
    ms, evals = timed(step_resident, K)
    sampler.mark_end()
    launches = clf.handle.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    sweep.reset()
    ms_e2e, evals_e2e = timed(run_e2e, K)
    sweep.reset()

    # roofline of the dominant kernel: event-bracket every conv launch over the same K steps
    roof = None
    if rank == 0:
        lib.fav_conv_timing_enable(h, 1)
        torch.cuda.synchronize()
        conv_ms, n_conv, exec_gflop = 0.0, 0, 0.0
        KR = min(K, 200)
        for i in range(KR):
            step_resident(i)
            msa, gfa, cnt = (C.c_float * 256)(), (C.c_float * 256)(), C.c_int()
            _lib.check(lib.fav_conv_timing_read_all(h, msa, gfa, 256, C.byref(cnt)), "timing")
            conv_ms += sum(msa[j] for j in range(cnt.value))
            exec_gflop += sum(gfa[j] for j in range(cnt.value))
            n_conv += cnt.value
        lib.fav_conv_timing_enable(h, 0)
        peaks = measured_peaks()
        flops = flops_per_eval(T_PASSES) * BLOCK * KR
        achieved = flops / (conv_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "conv_traffic.json")      # dram bytes per step from the committed ncu --set full capture
        if os.path.exists(tp):
            with open(tp) as fh:
                tj = json.load(fh)
                # bytes per image from the committed capture (taken at the step size named in the file; constant in the HBM-streaming
                # regime) x the images of one step here
                traffic = tj["dram_bytes_per_image"] * BLOCK if "dram_bytes_per_image" in tj else tj.get("dram_bytes_per_step")
        peak = peaks["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                "launches_per_step": n_conv / KR, "conv_ms_per_step": conv_ms / KR,
                "conv_share_of_step": conv_ms / KR / (ms / K),
                "flops_per_step": flops / KR,
                "executed_nominal_tflops": exec_gflop / conv_ms,
                "executed_note": "block 0 is pass-invariant and runs once per image (not T times), all-padding filter taps are "
                                 "skipped and 2x2-spatial convs are folded into dense GEMMs, so the MMAs actually issued are fewer "
                                 "than the SURVEY 8(d) algorithmic count used for 'achieved'; executed_nominal_tflops counts each "
                                 "launch's own 2*M*K*N",
                "note": "achieved = algorithmic FLOPs per step / summed conv-kernel device time per step "
                        "(CUDA events around each of the conv launches)"}
    sweep.reset()
    if world > 1:
        dist.barrier()

    if rank == 0:
        base = None if args.no_cpu_baseline else cpu_baseline()
        out = {
            "metric": "corrupted-image evals/sec", "value": evals / (ms * 1e-3), "unit": "evals/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, BLOCK),
            "clocks": clocks,
            "e2e": {"value": evals_e2e / (ms_e2e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": BLOCK * 32 * 32 * 3 + BLOCK * 4, "d2h_bytes_per_step": sweep.acc.words * 8},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": base,
            "tflops_whole_step": flops_per_eval(T_PASSES) * evals / (ms * 1e-3) / 1e12 / world,
        }
        emit_json(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
