"""CorruptionSweep -- the offline batch driver of the path (15 corruptions x 5 severities).

No reference counterpart: the closest thing is the playground's batch replay loop
(platform/backend/main.py:340-352), a scalar recurrence.  Results are emitted in the
reference's style -- plain dicts of rounded floats (trust_engine.py:247-263) and an in-memory
CSV like session_logger.py:15-51.

Sharding: work items (cell, image block) are dealt round-robin to ranks, weights
replicated, Philox counters keyed by the GLOBAL image index, and the only exchange is one
integer all-reduce of the histogram arena at the end -- so metrics are bit-identical on
1/2/4/8 GPUs.
"""
import csv
import ctypes as C
import io
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, spec
from .classifier import VisionClassifier, _ptr, _stream
from .spec import CorruptionConfig

HDR = 8


@dataclass
class SweepConfig:
    model: str = "resnet18"
    num_classes: int = 10
    input_hw: tuple = (32, 32)
    corruptions: tuple = spec.IMPLEMENTED
    severities: tuple = (1, 2, 3, 4, 5)
    include_clean: bool = False
    T: int = 20
    p_drop: float = 0.2
    tau: float = 0.9
    n_bins: int = 15
    n_buckets: int = 4096
    seed: int = 0
    weights_seed: int = 0
    logit_gain: float = None
    block: int = 512               # images per launch sequence

    def cells(self):
        out = [CorruptionConfig(None, 0)] if self.include_clean else []
        out += [CorruptionConfig(c, s) for c in self.corruptions for s in self.severities]
        return out

    def to_dict(self):
        d = dict(self.__dict__)
        d["corruptions"] = list(self.corruptions)
        d["severities"] = list(self.severities)
        d["input_hw"] = list(self.input_hw)
        return d


def partition(n_items, rank, world_size):
    """Indices of the work items owned by ``rank`` (round-robin: equal FLOPs, cells interleaved)."""
    return list(range(rank, n_items, world_size))


def allreduce_arena(arena):
    """The path's only exchange: integer sum of the histogram arena over ranks (NCCL on GPUs, gloo in the
    CPU tests).  Integer payload => the result is independent of reduction order and of the rank count."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(arena, op=dist.ReduceOp.SUM)
    return arena


def finalize(arena, num_classes, n_bins, n_buckets):
    """int64 arena of one cell -> metrics dict (fp64 on the host)."""
    a = np.asarray(arena, dtype=np.int64)
    n = int(a[0])
    out = {"n": n, "n_invalid": int(a[6])}          # labels outside [0, C) are counted, never histogrammed
    if n == 0:
        return out
    two32 = 4294967296.0
    out["accuracy"] = int(a[1]) / n
    out["failure_rate"] = int(a[2]) / n
    out["mean_confidence"] = int(a[3]) / two32 / n
    out["mean_entropy"] = int(a[4]) / two32 / n
    out["mean_mutual_information"] = int(a[5]) / two32 / n
    ece = 0.0
    for b in range(n_bins):
        cnt, sconf, ncor = (int(v) for v in a[HDR + 3 * b: HDR + 3 * b + 3])
        if cnt:
            ece += cnt / n * abs(ncor / cnt - sconf / two32 / cnt)
    out["ece"] = ece
    base = HDR + 3 * n_bins
    for s, name in enumerate(("auroc_msp", "auroc_entropy", "auroc_mi")):
        bk = a[base + 2 * s * n_buckets: base + 2 * (s + 1) * n_buckets].reshape(n_buckets, 2).astype(np.float64)
        neg, pos = bk[:, 0], bk[:, 1]
        P, Nn = pos.sum(), neg.sum()
        if P == 0 or Nn == 0:
            out[name] = float("nan")
        else:
            below = np.concatenate([[0.0], np.cumsum(neg)[:-1]])
            out[name] = float((pos * (below + 0.5 * neg)).sum() / (P * Nn))
    return out


class MetricsAccumulator:
    """Device arena of integer histograms, one row per sweep cell."""

    def __init__(self, clf: VisionClassifier, n_cells, n_bins=15, n_buckets=4096):
        self.clf, self.n_bins, self.n_buckets = clf, n_bins, n_buckets
        self.words = int(clf.lib.fav_hist_words(clf.num_classes, n_bins, n_buckets))
        self.arena = torch.zeros((n_cells, self.words), dtype=torch.int64, device=clf.device)

    def reset(self):
        self.arena.zero_()

    def add_logits(self, cell, logits, labels, tau, outputs=None):
        n, T, c = logits.shape
        o = outputs or {}
        _lib.check(self.clf.lib.fav_epilogue_accumulate(
            self.clf.handle.h, _ptr(logits), _ptr(labels), n, T, c, float(tau), self.n_bins, self.n_buckets,
            C.c_void_p(self.arena[cell].data_ptr()), _ptr(o.get("confidence")), _ptr(o.get("entropy")),
            _ptr(o.get("mutual_information")), _ptr(o.get("pred")), _ptr(o.get("failure_flag")), _stream(self.clf.device)),
            "fav_epilogue_accumulate")

    def add_scores(self, cell, conf, ent, mi, pred, labels, tau):
        _lib.check(self.clf.lib.fav_accumulate(
            self.clf.handle.h, _ptr(conf), _ptr(ent), _ptr(mi), _ptr(pred), _ptr(labels), conf.numel(),
            self.clf.num_classes, float(tau), self.n_bins, self.n_buckets, C.c_void_p(self.arena[cell].data_ptr()),
            _stream(self.clf.device)), "fav_accumulate")

    def allreduce(self):
        """The path's only exchange: integer sum of the arena over the ranks.  On GPUs this is the library's own
        fav_allreduce (ncclAllReduce(sum, int64) on the current stream; the communicator is created on first use, its id
        broadcast through torch.distributed); FAV_ALLREDUCE=torch or a CPU arena (gloo tests) uses torch.distributed."""
        import os
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if not self.arena.is_cuda or os.environ.get("FAV_ALLREDUCE", "nccl") == "torch":
            allreduce_arena(self.arena)
            return
        self.init_comm()
        handle = self.clf.handle
        _lib.check(self.clf.lib.fav_allreduce(handle.h, _ptr(self.arena), self.arena.numel(), _stream(self.clf.device)), "fav_allreduce")

    def init_comm(self):
        """Create the library's NCCL communicator (collective; ~0.5 s once per process): rank 0's unique id travels
        through torch.distributed.  Called by CorruptionSweep.prepare() so that it stays out of timed regions."""
        import torch.distributed as dist
        handle = self.clf.handle
        if getattr(handle, "comm_ready", False) or not self.arena.is_cuda:
            return
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        uid = torch.zeros(128, dtype=torch.uint8, device=self.arena.device)
        if dist.get_rank() == 0:
            buf = C.create_string_buffer(128)
            _lib.check(self.clf.lib.fav_comm_unique_id(buf), "fav_comm_unique_id")
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        raw = bytes(uid.cpu().numpy().tobytes())
        _lib.check(self.clf.lib.fav_comm_init(handle.h, C.c_char_p(raw), dist.get_rank(), dist.get_world_size()), "fav_comm_init")
        handle.comm_ready = True
        warm = torch.zeros(8, dtype=torch.int64, device=self.arena.device)          # first collective sets up the channels
        _lib.check(self.clf.lib.fav_allreduce(handle.h, _ptr(warm), warm.numel(), _stream(self.clf.device)), "fav_allreduce")
        torch.cuda.current_stream(self.arena.device).synchronize()

    def results(self):
        if getattr(self, "_host", None) is None or self._host.shape != self.arena.shape:
            self._host = torch.empty(self.arena.shape, dtype=torch.int64).pin_memory()
        self._host.copy_(self.arena, non_blocking=True)
        torch.cuda.current_stream(self.arena.device).synchronize()
        host = self._host.numpy()
        return [finalize(host[i], self.clf.num_classes, self.n_bins, self.n_buckets) for i in range(host.shape[0])]


class CorruptionSweep:
    def __init__(self, cfg: SweepConfig = None, classifier: VisionClassifier = None, device=0):
        self.cfg = cfg or SweepConfig()
        self.clf = classifier or VisionClassifier(self.cfg.model, self.cfg.num_classes, self.cfg.input_hw,
                                                  self.cfg.weights_seed, self.cfg.logit_gain, device=device)
        self.cells = self.cfg.cells()
        self.acc = MetricsAccumulator(self.clf, len(self.cells), self.cfg.n_bins, self.cfg.n_buckets)
        self._x = None
        self._logits = None
        self._stream_state = None

    def reset(self):
        self.acc.reset()

    def prepare(self, n_images=None):
        """Build every cell's host table, upload it, size the workspace and touch every kernel once (untimed
        setup: Poisson / stencil tables take tens of ms each on the host)."""
        n = min(self.cfg.block, n_images or self.cfg.block)
        h, w = self.cfg.input_hw
        # a FULL block: the library grows its K1 scratch (up to 7 fp32 planes per image for elastic_transform) and builds
        # every cell's table on first use -- both synchronise, so they must happen here and not inside a timed step
        x = torch.zeros((n, h, w, 3), dtype=torch.uint8, device=self.clf.device)
        for cell in self.cells:
            self.clf.corrupt_normalize(x, cell, self.cfg.seed, 0)
        _lib.check(self.clf.lib.fav_reserve(self.clf.handle.h, n, self.cfg.T), "fav_reserve")
        self._buffers(n)
        import os
        if os.environ.get("FAV_ALLREDUCE", "nccl") != "torch":
            self.acc.init_comm()
        torch.cuda.synchronize()

    def cell_order(self):
        """Cells in a strided order: with the grid laid out corruption-major, consecutive steps visit different corruptions
        AND different severities (stride coprime to the cell count, ~ one fifth of it), so any short run of steps -- a
        bench window, one rank's round-robin share -- samples the whole grid instead of its first few corruptions."""
        n = len(self.cells)
        stride = max(1, n // 5 + 1)
        while math.gcd(stride, n) != 1:
            stride += 1
        return [(k * stride) % n for k in range(n)]

    def work_items(self, n_images):
        nblk = (n_images + self.cfg.block - 1) // self.cfg.block
        order = self.cell_order()
        return [(ci, b) for b in range(nblk) for ci in order]

    def _buffers(self, n):
        h, w = self.cfg.input_hw
        if self._x is None or self._x.shape[0] < n:
            self._x = torch.empty((n, h, w, 3), dtype=torch.bfloat16, device=self.clf.device)
            self._logits = torch.empty((n, self.cfg.T, self.cfg.num_classes), dtype=torch.float32, device=self.clf.device)
        return self._x[:n], self._logits[:n]

    def run_item(self, images_dev, labels_dev, item, first_image=0):
        """One step of the hot path: one image block of one (corruption, severity) cell."""
        ci, b = item
        lo = b * self.cfg.block
        hi = min(lo + self.cfg.block, images_dev.shape[0])
        x, logits = self._buffers(hi - lo)
        cfg = self.cfg
        self.clf.corrupt_normalize(images_dev[lo:hi], self.cells[ci], cfg.seed, first_image + lo, out=x)
        self.clf.forward_logits(x, cfg.T, cfg.p_drop, cfg.seed, first_image + lo, out=logits)
        self.acc.add_logits(ci, logits, labels_dev[lo:hi], cfg.tau)
        return hi - lo

    def _pipe_state(self):
        """Two staging slots for the software-pipelined paths: uint8 / label staging (host data only), the bf16 K1 output, a
        pinned arena row, and the events that order the side stream (copies + K1) against the main stream (forward + K3/K4)."""
        cfg, dev = self.cfg, self.clf.device
        h, w = cfg.input_hw
        if self._stream_state is None:
            B = cfg.block
            self._stream_state = dict(
                side=torch.cuda.Stream(dev),
                img=None, lab=None,
                x=[torch.empty((B, h, w, 3), dtype=torch.bfloat16, device=dev) for _ in range(2)],
                row=[torch.empty(self.acc.words, dtype=torch.int64).pin_memory() for _ in range(2)],
                ready=[torch.cuda.Event() for _ in range(2)], done=[torch.cuda.Event() for _ in range(2)])
        return self._stream_state

    def _pipeline(self, items, n_img, fetch, first_image, on_row=None):
        """Software pipeline over `items`: on the SIDE stream step k+1's input is fetched (`fetch(slot, lo, hi)` -> uint8
        device block + int32 device labels: an H2D copy for host data, a view for resident data) and taken through K1
        (corrupt + normalize) while the MAIN stream runs step k's forward and K3+K4.  The conv kernels fill every SM, so
        the K1 CTAs run in the ramp-up / drain gaps between the conv launches instead of adding to the step.  Same kernels,
        same inputs, same order of the integer accumulation per cell: results are bit-identical to the sequential path."""
        cfg, dev = self.cfg, self.clf.device
        ss = self._pipe_state()
        cur, side = torch.cuda.current_stream(dev), ss["side"]
        if ss.get("logits") is None:
            ss["logits"] = torch.empty((cfg.block, cfg.T, cfg.num_classes), dtype=torch.float32, device=dev)
        labs = [None, None]

        def span(item):
            lo = item[1] * cfg.block
            return lo, min(lo + cfg.block, n_img)

        def stage(k):
            slot = k & 1
            lo, hi = span(items[k])
            with torch.cuda.stream(side):
                if k >= 2:
                    side.wait_event(ss["done"][slot])            # step k-2 is done with this slot's buffers
                else:
                    side.wait_stream(cur)
                img, labs[slot] = fetch(slot, lo, hi)
                self.clf.corrupt_normalize(img, self.cells[items[k][0]], cfg.seed, first_image + lo, out=ss["x"][slot][:hi - lo])
                ss["ready"][slot].record(side)

        evals = 0
        if len(items):
            stage(0)
        for k, item in enumerate(items):
            slot = k & 1
            if k + 1 < len(items):
                stage(k + 1)
            lo, hi = span(item)
            n = hi - lo
            cur.wait_event(ss["ready"][slot])
            logits = ss["logits"][:n]
            self.clf.forward_logits(ss["x"][slot][:n], cfg.T, cfg.p_drop, cfg.seed, first_image + lo, out=logits)
            self.acc.add_logits(item[0], logits, labs[slot], cfg.tau)
            if on_row is not None:
                ss["row"][slot].copy_(self.acc.arena[item[0]], non_blocking=True)
            ss["done"][slot].record(cur)
            if on_row is not None and k >= 1:
                ss["done"][slot ^ 1].synchronize()
                on_row(items[k - 1], ss["row"][slot ^ 1])
            evals += n
        if len(items) and on_row is not None:
            last = (len(items) - 1) & 1
            ss["done"][last].synchronize()
            on_row(items[-1], ss["row"][last])
        side.wait_stream(cur)                                    # leave both streams ordered for whoever comes next
        cur.wait_stream(side)
        return evals

    def run_items(self, images_dev, labels_dev, items, first_image=0):
        """The steps `items` on RESIDENT device data, software-pipelined (K1 of step k+1 beside the forward of step k)."""
        return self._pipeline(items, images_dev.shape[0], lambda slot, lo, hi: (images_dev[lo:hi], labels_dev[lo:hi]), first_image)

    def run_stream(self, host_images, host_labels, items, first_image=0, on_row=None):
        """The steps `items` on HOST blocks (uint8 [N,H,W,3] / int32 [N] torch tensors, pinned for async copies): block k+1 is
        copied host->device and corrupted on a side stream while block k's forward runs, and after every step the cell's
        histogram-arena row is read back into pinned memory; `on_row(item, row)` sees it one step later (the only host
        synchronisation).  Returns the number of evals."""
        cfg, dev = self.cfg, self.clf.device
        h, w = cfg.input_hw
        ss = self._pipe_state()
        if ss["img"] is None:
            ss["img"] = [torch.empty((cfg.block, h, w, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
            ss["lab"] = [torch.empty(cfg.block, dtype=torch.int32, device=dev) for _ in range(2)]

        def fetch(slot, lo, hi):
            ss["img"][slot][:hi - lo].copy_(host_images[lo:hi], non_blocking=True)
            ss["lab"][slot][:hi - lo].copy_(host_labels[lo:hi], non_blocking=True)
            return ss["img"][slot][:hi - lo], ss["lab"][slot][:hi - lo]

        return self._pipeline(items, host_images.shape[0], fetch, first_image, on_row if on_row is not None else (lambda item, row: None))

    def run(self, images_u8, labels, rank=0, world_size=1, first_image=0, timing=False):
        """images uint8 [N,H,W,3] (numpy or torch, host or device), labels int [N].
        Returns {(corruption, severity): metrics} after the cross-rank reduction.  Starts from an empty arena (after the
        all-reduce every rank holds the global sums: a second run must not add to them); run_item / run_stream are the
        accumulate-only calls."""
        import numpy as np
        self.acc.reset()
        if isinstance(images_u8, np.ndarray):
            images_u8 = torch.from_numpy(np.ascontiguousarray(images_u8))
        if isinstance(labels, np.ndarray):
            labels = torch.from_numpy(np.ascontiguousarray(labels))
        items = self.work_items(images_u8.shape[0])
        mine = [items[i] for i in partition(len(items), rank, world_size)]
        if not images_u8.is_cuda:
            # host data: stream the blocks through two pinned-to-device staging slots, copies overlapped with compute
            self.clf._images(images_u8[:0])                  # shape / dtype validation
            host_images = images_u8.contiguous().pin_memory()
            host_labels = torch.as_tensor(labels).to(torch.int32).contiguous().pin_memory()
            self.run_stream(host_images, host_labels, mine, first_image)
        else:
            images_dev = self.clf._images(images_u8)
            labels_dev = self.clf._labels(labels)
            evs = []
            if not timing:
                self.run_items(images_dev, labels_dev, mine, first_image)
            for item in (mine if timing else []):
                if timing:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(torch.cuda.current_stream(self.clf.device))
                n = self.run_item(images_dev, labels_dev, item, first_image)
                if timing:
                    e1.record(torch.cuda.current_stream(self.clf.device))
                    evs.append((item[0], n, e0, e1))
            if timing:                                   # device ms and evals per cell on this rank
                torch.cuda.current_stream(self.clf.device).synchronize()
                self.cell_ms = [0.0] * len(self.cells)
                self.cell_evals = [0] * len(self.cells)
                for ci, n, e0, e1 in evs:
                    self.cell_ms[ci] += e0.elapsed_time(e1)
                    self.cell_evals[ci] += n
        self.acc.allreduce()
        res = self.acc.results()
        return {(c.name or "clean", c.severity): r for c, r in zip(self.cells, res)}

    COLUMNS = ["corruption", "severity", "n", "accuracy", "ece", "mean_confidence", "mean_entropy",
               "mean_mutual_information", "failure_rate", "auroc_msp", "auroc_entropy", "auroc_mi"]
    PERF_COLUMNS = ["gpu_ms", "evals_per_gpu_s", "tflops", "roofline_frac"]

    @staticmethod
    def records(results, perf=None):
        """Per-cell records in the reference's style: plain dicts of rounded floats (trust_engine.py:247-263), NaN -> None so
        that json.dumps emits valid JSON like the websocket payload at main.py:200.  perf: optional {cell: {gpu_ms, ...}}
        (the sweep's metrics rows also carry throughput and roofline fraction)."""
        out = []
        for (name, sev), r in results.items():
            rec = {"corruption": name, "severity": sev}
            for k in CorruptionSweep.COLUMNS[2:]:
                v = r.get(k)
                rec[k] = None if (isinstance(v, float) and math.isnan(v)) else (round(v, 6) if isinstance(v, float) else v)
            if perf is not None and (name, sev) in perf:
                for k in CorruptionSweep.PERF_COLUMNS:
                    v = perf[(name, sev)].get(k)
                    rec[k] = round(v, 6) if isinstance(v, float) else v
            out.append(rec)
        return out

    @staticmethod
    def to_csv(results, perf=None):
        """In-memory CSV in the style of session_logger.py:15-51 (header row, one row per record)."""
        cols = CorruptionSweep.COLUMNS + (CorruptionSweep.PERF_COLUMNS if perf is not None else [])
        buf = io.StringIO()
        wr = csv.writer(buf)
        wr.writerow(cols)
        for rec in CorruptionSweep.records(results, perf):
            wr.writerow(["" if rec.get(k) is None else rec.get(k) for k in cols])
        return buf.getvalue()

    @staticmethod
    def to_json(results, perf=None, header=None):
        """One JSON document: {'type': 'sweep_result', 'config': ..., 'cells': [records]} -- the shape of the reference's
        batch reply {'type': 'sequence_result', 'data': [...]} (main.py:354-357)."""
        import json
        doc = {"type": "sweep_result"}
        doc.update(header or {})
        doc["cells"] = CorruptionSweep.records(results, perf)
        return json.dumps(doc)


def measured_peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json at the repo root when present, else the profiling
    guide's fallback numbers."""
    import json
    import os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def step_gflop(sweep, images_dev, labels_dev):
    """Nominal conv GFLOP of ONE block through the classifier (each launch's own 2*M*K*N as recorded by the library's
    per-launch timing: the pass-invariant prefix once per image, the rest T times)."""
    lib, h = sweep.clf.lib, sweep.clf.handle.h
    lib.fav_conv_timing_enable(h, 1)
    sweep.run_item(images_dev, labels_dev, (0, 0))
    ms, gf, cnt = (C.c_float * 512)(), (C.c_float * 512)(), C.c_int()
    _lib.check(lib.fav_conv_timing_read_all(h, ms, gf, 512, C.byref(cnt)), "fav_conv_timing_read_all")
    lib.fav_conv_timing_enable(h, 0)
    return float(sum(gf[i] for i in range(cnt.value)))


def main(argv=None):
    """``python -m fav.sweep``: a corruption sweep on synthetic Philox images (there are no datasets offline).  Results per
    (corruption, severity) cell -- n, accuracy, ECE, mean confidence / entropy / MI, failure rate, AUROC x 3 plus device time,
    evals/s and roofline fraction -- as CSV (session_logger.py style) or one JSON document on stdout or --out.  Under
    torchrun every rank takes its share of the (cell, block) items; rank 0 prints."""
    import argparse
    import ctypes
    import os
    import sys
    import torch.distributed as dist
    ap = argparse.ArgumentParser(prog="python -m fav.sweep", description=main.__doc__)
    ap.add_argument("--model", default="resnet18", choices=["resnet18", "resnet50"])
    ap.add_argument("--hw", type=int, default=32, help="square input size (32: CIFAR profile, 224: ImageNet profile)")
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--images", type=int, default=1000)
    ap.add_argument("--passes", type=int, default=20, help="MC-dropout passes T (1 = deterministic MSP)")
    ap.add_argument("--p-drop", type=float, default=0.2)
    ap.add_argument("--tau", type=float, default=0.9)
    ap.add_argument("--block", type=int, default=2048)
    ap.add_argument("--corruptions", default="all", help="comma-separated names or 'all'")
    ap.add_argument("--severities", default="1,2,3,4,5")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--logit-gain", type=float, default=8.0, help="fixture for random-init weights (SURVEY.md section 7)")
    ap.add_argument("--format", default="csv", choices=["csv", "json"])
    ap.add_argument("--out", default="-", help="output file ('-' = stdout)")
    a = ap.parse_args(argv)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    names = tuple(spec.IMPLEMENTED) if a.corruptions == "all" else tuple(x for x in a.corruptions.split(",") if x)
    cfg = SweepConfig(model=a.model, num_classes=a.classes, input_hw=(a.hw, a.hw), T=a.passes, p_drop=a.p_drop, tau=a.tau,
                      block=min(a.block, a.images), seed=a.seed, logit_gain=a.logit_gain, corruptions=names,
                      severities=tuple(int(x) for x in a.severities.split(",")))
    sw = CorruptionSweep(cfg, device=local)
    sw.prepare(a.images)
    dev = sw.clf.device
    x = torch.empty((a.images, a.hw, a.hw, 3), dtype=torch.uint8, device=dev)
    y = torch.empty(a.images, dtype=torch.int32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(sw.clf.lib.fav_synth_images(sw.clf.handle.h, _ptr(x), a.images, a.hw, a.hw, a.seed, 0, st), "fav_synth_images")
    _lib.check(sw.clf.lib.fav_synth_labels(sw.clf.handle.h, _ptr(y), a.images, a.classes, a.seed, 0, st), "fav_synth_labels")
    gflop_block = step_gflop(sw, x, y)                       # also warms every kernel of the forward
    res = sw.run(x, y, rank=rank, world_size=world, timing=True)
    ms = torch.tensor(sw.cell_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.SUM)            # GPU-milliseconds per cell, summed over the ranks
    if rank == 0:
        peaks = measured_peaks()
        gflop_eval = gflop_block / min(cfg.block, a.images)
        perf = {}
        for c, r, m in zip(sw.cells, res.values(), ms.tolist()):
            tf = gflop_eval * r["n"] / m if m > 0 else 0.0                   # GFLOP / ms = TFLOP/s (per GPU)
            perf[(c.name or "clean", c.severity)] = {"gpu_ms": m, "evals_per_gpu_s": r["n"] / m * 1e3 if m > 0 else 0.0,
                                                      "tflops": tf, "roofline_frac": tf / peaks["bf16_tflops_sustained"]}
        if a.format == "csv":
            text = CorruptionSweep.to_csv(res, perf)
        else:
            text = CorruptionSweep.to_json(res, perf, {"config": cfg.to_dict(), "images": a.images, "n_gpus": world,
                                                      "gflop_per_eval": gflop_eval, "roofline_peak_tflops": peaks["bf16_tflops_sustained"],
                                                      "roofline_peak_source": peaks["source"] + " (sustained cuBLAS bf16)"}) + "\n"
        if a.out == "-":
            sys.stdout.write(text)
        else:
            with open(a.out, "w") as fh:
                fh.write(text)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
