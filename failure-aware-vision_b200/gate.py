"""UncertaintyGate -- streaming per-frame provider with SignalAnalyzer's exact contract.

Substitutes ``SignalAnalyzer.analyze_frame`` at platform/backend/main.py:160 (class at
signal_analyzer.py:18-171): same constants (:22-34), same return-dict keys (:126-142), same
status rules (:145-171) -- so ``TrustEngine.update(vision_status, anomaly_score, dt)``
(trust_engine.py:139) keeps working unchanged.  The pixel arithmetic (gray, Laplacian sums,
brightness, frame difference, histogram; :62-105) runs in one fused integer CUDA kernel
(``fav_frame_stats``) that is bit-exact against OpenCV; the classifier's uncertainty is
computed from the same device frame and reported under ``metrics['uncertainty']``.

``score_source``: 'uncertainty' (default) -> anomaly_score = clip(H / ln C, 0, 1);
'signal' -> the reference's fused four-metric score (used for golden parity);
'max' -> the larger of the two.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .classifier import VisionClassifier, _ptr, _stream


class SignalFinisher:
    """Host half of SignalAnalyzer.analyze_frame: turns the integer frame statistics
    {sum_lap, sum_lap_sq, sum_gray, sum_absdiff, hist[256]} into the reference's scores and
    status, line by line after signal_analyzer.py:65-123,145-171 (pure Python / numpy, fp64;
    the histogram entropy keeps the reference's float32 numpy expressions)."""

    # fusion weights and thresholds: signal_analyzer.py:22-34
    W_BLUR, W_BRIGHTNESS, W_FREEZE, W_ENTROPY = 0.35, 0.25, 0.15, 0.25
    FREEZE_DIFF_THRESHOLD = 1.0
    FREEZE_CONSEC_NEEDED = 5
    BLANK_BRIGHTNESS_LO, BLANK_BRIGHTNESS_HI = 15, 245
    CORRUPT_ENTROPY_LO, CORRUPT_ENTROPY_HI = 2.0, 7.5
    BLUR_BASELINE = 500.0

    def __init__(self):
        self.reset()

    def reset(self):
        self.have_prev = False
        self.consecutive_frozen = 0

    def finish(self, st, n):
        s_lap, s_lap2, s_gray, s_diff = (int(v) for v in st[:4])
        hist = np.asarray(st[4:260])
        # 1. blur (:65-67) -- exact variance from integer sums
        laplacian_var = (n * s_lap2 - s_lap * s_lap) / (n * n)
        blur_score = max(0.0, min(1.0, 1.0 - laplacian_var / self.BLUR_BASELINE))
        # 2. brightness (:70-73)
        mean_brightness = s_gray / n
        brightness_score = max(0.0, min(1.0, abs(mean_brightness - 128.0) / 128.0))
        # 3. freeze (:76-96)
        if self.have_prev:
            mean_diff = s_diff / n
            if mean_diff < self.FREEZE_DIFF_THRESHOLD:
                self.consecutive_frozen += 1
            else:
                self.consecutive_frozen = 0
            if self.consecutive_frozen >= self.FREEZE_CONSEC_NEEDED:
                freeze_score = 1.0
            elif self.consecutive_frozen > 0:
                freeze_score = 0.3 * (self.consecutive_frozen / self.FREEZE_CONSEC_NEEDED)
            else:
                freeze_score = 0.0
        else:
            freeze_score, mean_diff = 0.0, 10.0
        self.have_prev = True
        # 4. entropy (:101-112)
        histogram = hist.astype(np.float32)
        histogram = histogram / (histogram.sum() + 1e-10)
        histogram = histogram[histogram > 0]
        entropy = float(-np.sum(histogram * np.log2(histogram)))
        if entropy < 4.0:
            entropy_score = max(0.0, min(1.0, (4.0 - entropy) / 4.0))
        elif entropy > 7.0:
            entropy_score = max(0.0, min(1.0, (entropy - 7.0) / 1.5))
        else:
            entropy_score = 0.0
        signal_score = (self.W_BLUR * blur_score + self.W_BRIGHTNESS * brightness_score
                        + self.W_FREEZE * freeze_score + self.W_ENTROPY * entropy_score)
        signal_score = max(0.0, min(1.0, signal_score))
        # status priority rules (:145-171)
        if mean_brightness < self.BLANK_BRIGHTNESS_LO or mean_brightness > self.BLANK_BRIGHTNESS_HI:
            status = 'VISION_BLANK'
        elif self.consecutive_frozen >= self.FREEZE_CONSEC_NEEDED:
            status = 'VISION_FROZEN'
        elif entropy < self.CORRUPT_ENTROPY_LO or entropy > self.CORRUPT_ENTROPY_HI:
            status = 'VISION_CORRUPTED'
        else:
            status = 'VISION_OK'
        return dict(laplacian_var=laplacian_var, blur_score=blur_score, mean_brightness=mean_brightness,
                    brightness_score=brightness_score, mean_diff=mean_diff, freeze_score=freeze_score,
                    entropy=entropy, entropy_score=entropy_score, signal_score=signal_score, vision_status=status)


def assemble_result(fin, unc, score_source="uncertainty", tau=0.9, num_classes=1000):
    """The dict SignalAnalyzer.analyze_frame returns (signal_analyzer.py:126-142: same keys, same rounding), with the
    classifier's uncertainty under metrics['uncertainty'] and, depending on score_source, in anomaly_score.
    fin: SignalFinisher.finish(...); unc: (confidence, entropy, mutual_information, pred) or None.  Pure host code."""
    signal_score, vision_status = fin["signal_score"], fin["vision_status"]
    anomaly_score, u = signal_score, None
    if unc is not None:
        conf, H, mi, pred = unc
        norm_h = max(0.0, min(1.0, H / math.log(num_classes)))
        u = {"confidence": round(conf, 6), "entropy": round(H, 6), "mutual_information": round(mi, 6),
             "pred": int(pred), "normalized_entropy": round(norm_h, 6), "high_confidence": bool(conf >= tau)}
        if score_source == "uncertainty":
            anomaly_score = norm_h
        elif score_source == "max":
            anomaly_score = max(signal_score, norm_h)
    out = {
        'anomaly_score': round(anomaly_score, 6),
        'vision_status': vision_status,
        'metrics': {
            'blur': round(fin["blur_score"], 4),
            'brightness': round(fin["brightness_score"], 4),
            'freeze': round(fin["freeze_score"], 4),
            'entropy': round(fin["entropy_score"], 4),
            'raw': {
                'laplacian_var': round(fin["laplacian_var"], 2),
                'mean_brightness': round(fin["mean_brightness"], 1),
                'frame_diff': round(fin["mean_diff"], 2),
                'entropy': round(fin["entropy"], 3),
            }
        }
    }
    if u is not None:
        out['metrics']['uncertainty'] = u
    return out


class UncertaintyGate:
    def __init__(self, classifier: VisionClassifier = None, frame_hw=(480, 640), T=1, p_drop=0.2, tau=0.9,
                 score_source="uncertainty", model="resnet18", num_classes=1000, device=0, weights_seed=0,
                 logit_gain=None, use_classifier=True, use_graph=None, mask_schedule="per_frame"):
        if not torch.cuda.is_available():
            raise RuntimeError("UncertaintyGate needs a CUDA device (sm_100a); there is no CPU fallback")
        self.frame_hw = tuple(frame_hw)
        self.T, self.p_drop, self.tau, self.score_source = int(T), float(p_drop), float(tau), score_source
        self.clf = classifier
        if self.clf is None and use_classifier:
            self.clf = VisionClassifier(model, num_classes, self.frame_hw, weights_seed, logit_gain, device=device)
        self.device = self.clf.device if self.clf is not None else torch.device("cuda", device)
        self.handle = self.clf.handle if self.clf is not None else _lib.Handle(device)
        self.lib = self.handle.lib
        h, w = self.frame_hw
        self._pinned = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        self._frame = torch.empty((1, h, w, 3), dtype=torch.uint8, device=self.device)
        self._gray = torch.zeros((h, w), dtype=torch.uint8, device=self.device)
        self._stats = torch.zeros(260, dtype=torch.int64, device=self.device)
        self._stats_host = torch.empty(260, dtype=torch.int64).pin_memory()
        self._packed_host = torch.empty((1, 4), dtype=torch.float32).pin_memory()
        self._side = torch.cuda.Stream(self.device)
        self._fin = SignalFinisher()
        # CUDA-graph replay of the whole device side of a frame (H2D copy, frame statistics, K1..K3, D2H copies): one
        # launch instead of ~30.  The dropout masks are keyed by first_image, a kernel argument frozen at capture:
        #   mask_schedule = "per_frame" (default): frame i uses masks keyed by i -> eager launches when T > 1;
        #   mask_schedule = "frozen": every frame sees the SAME T masks (common random numbers: the score is a
        #                   deterministic function of the frame, differences between frames carry no mask noise) -> graph.
        if mask_schedule not in ("per_frame", "frozen"):
            raise ValueError("mask_schedule must be 'per_frame' or 'frozen'")
        self.mask_schedule = mask_schedule
        self.use_graph = (self.T == 1 or mask_schedule == "frozen") if use_graph is None else bool(use_graph)
        self._graph = None
        self._graph_failed = False
        self.graph_error = None
        self.reset()

    @property
    def graph_active(self):
        """True once the per-frame device work replays as one CUDA graph (the latency harness asserts on it)."""
        return self._graph is not None

    def _enqueue_frame(self, first, first_image):
        """Device side of one frame on the current stream: pinned frame -> device, fused SignalAnalyzer statistics,
        classifier + uncertainty epilogue, results -> pinned host buffers.  No host synchronisation."""
        h, w = self.frame_hw
        cur = torch.cuda.current_stream(self.device)
        self._frame[0].copy_(self._pinned, non_blocking=True)
        # the SignalAnalyzer statistics (one kernel) run beside the classifier chain on a side stream (fork / join; under
        # CUDA-graph capture this becomes a parallel branch of the graph)
        side = self._side if self.clf is not None else cur
        if side is not cur:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            _lib.check(self.lib.fav_frame_stats(self.handle.h, _ptr(self._frame), _ptr(self._gray), h, w,
                                                1 if first else 0, _ptr(self._stats), _stream(self.device)), "fav_frame_stats")
            self._stats_host.copy_(self._stats, non_blocking=True)
        if self.clf is not None:
            u = self.clf.uncertainty(self._frame, None, T=self.T, p=self.p_drop, seed=0, first_image=first_image, bgr=True)
            packed = torch.stack([u["confidence"], u["entropy"], u["mutual_information"], u["pred"].float()], 1)
            self._packed_host.copy_(packed, non_blocking=True)
        if side is not cur:
            cur.wait_stream(side)

    def _capture(self):
        try:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_frame(False, 0)
            self._graph = g
        except Exception as e:                 # capture is an optimisation: a failure falls back to eager launches, LOUDLY
            import warnings
            self._graph = None
            self._graph_failed = True
            self.graph_error = repr(e)
            warnings.warn(f"UncertaintyGate: CUDA-graph capture failed, running ~30 eager launches per frame instead: {e!r}",
                          RuntimeWarning)
            torch.cuda.synchronize(self.device)

    def reset(self):
        """Clear internal state (signal_analyzer.py:41-45)."""
        self._fin.reset()
        self._frame_count = 0

    # ------------------------------------------------------------------
    def analyze_frame(self, frame):
        """frame: BGR uint8 numpy [H,W,3] (video_source.py:105-117 hands out exactly this)."""
        h, w = self.frame_hw
        if frame.shape != (h, w, 3) or frame.dtype != np.uint8:
            raise ValueError(f"frame must be uint8 {(h, w, 3)}, got {frame.dtype} {frame.shape}")
        self._frame_count += 1
        self._pinned.copy_(torch.from_numpy(frame))
        first = not self._fin.have_prev
        if self.use_graph and not first and not self._graph_failed and self._frame_count > 2:
            if self._graph is None:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
                torch.cuda.current_stream(self.device).synchronize()
        if self._graph is None or first or self._frame_count <= 2:
            self._enqueue_frame(first, 0 if self.mask_schedule == "frozen" else self._frame_count)
            torch.cuda.current_stream(self.device).synchronize()
        fin = self._fin.finish(self._stats_host.numpy(), h * w)
        unc = tuple(float(v) for v in self._packed_host[0]) if self.clf is not None else None
        return assemble_result(fin, unc, self.score_source, self.tau, self.clf.num_classes if self.clf is not None else 0)
