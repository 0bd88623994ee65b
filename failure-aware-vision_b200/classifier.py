"""VisionClassifier -- the predict / uncertainty entry points of the path.

Reference seam: the per-frame provider called at platform/backend/main.py:160
(``analyzer.analyze_frame(frame)``) and the ML-score slot of
``AnomalySimulator.compute_anomaly`` (anomaly_simulator.py:34-77, "No PyTorch dependency").
The reference has no classifier; this class is the B200-native provider: corrupt -> normalize
-> ResNet x T MC-dropout passes -> softmax / confidence / entropy / mutual information ->
failure flag (README.md:22-24).  All compute runs in libfav_b200.so (hand-written sm_100a
kernels) through the C ABI; torch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, spec, weights
from .spec import CorruptionConfig


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device=None):
    """The caller's current stream ON `device` (a handle's kernels must run on the handle's own device)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class VisionClassifier:
    """ResNet-18/50 classifier with MC-dropout uncertainty on one GPU.

    Plain class with ``reset()`` like every per-connection object of the reference
    (signal_analyzer.py:41-45, trust_engine.py:37)."""

    def __init__(self, model="resnet18", num_classes=10, input_hw=(32, 32), weights_seed=0, logit_gain=None,
                 net=None, device=0, mean_std=None, profile=None):
        if not torch.cuda.is_available():
            raise RuntimeError("VisionClassifier needs a CUDA device (sm_100a); there is no CPU fallback")
        self.model, self.num_classes, self.input_hw = model, int(num_classes), tuple(input_hw)
        self.device = torch.device("cuda", device)          # the process-wide current device is left alone
        self.handle = _lib.Handle(device)
        self.lib = self.handle.lib
        self.profile = profile or spec.profile_for(*self.input_hw)
        self.mean, self.std = mean_std or spec.MEAN_STD[self.profile]
        self.net = net if net is not None else weights.build_model(model, num_classes, weights_seed, logit_gain)
        blob = weights.pack_resnet(self.net, model)
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(self.lib.fav_load_weights(self.handle.h, buf, len(blob), weights.MODEL_IDS[model], self.num_classes,
                                             self.input_hw[0], self.input_hw[1]), "fav_load_weights")

    # ------------------------------------------------------------------ housekeeping
    def reset(self):
        self.handle.reset()

    def _st(self):
        return _stream(self.device)

    def _images(self, images_u8):
        if isinstance(images_u8, np.ndarray):
            images_u8 = torch.from_numpy(np.ascontiguousarray(images_u8))
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError("images must be uint8 [N,H,W,3]")
        if tuple(images_u8.shape[1:3]) != self.input_hw:
            raise ValueError(f"images are {tuple(images_u8.shape[1:3])}, classifier was built for {self.input_hw}")
        if not images_u8.is_cuda:
            images_u8 = images_u8.pin_memory().to(self.device, non_blocking=True)
        return images_u8.contiguous()

    # ------------------------------------------------------------------ K1
    def corrupt_normalize(self, images_u8, corruption=None, seed=0, first_image=0, bgr=False, out=None,
                          out_f32=False, normalize=True):
        """uint8 [N,H,W,3] -> bf16 (or fp32) [N,H,W,3], corrupted then normalised, RGB order."""
        cfg = corruption if isinstance(corruption, CorruptionConfig) else CorruptionConfig(*(corruption or (None, 0)))
        x = self._images(images_u8)
        n, h, w, _ = x.shape
        if out is None:
            out = torch.empty((n, h, w, 3), dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
        # every per-cell constant / table is derived inside the library (csrc/tables.cu) and cached in the handle
        flags = (1 if bgr else 0) | (2 if out_f32 else 0) | (0 if normalize else 4) | _lib.PROFILE_FLAGS[self.profile]
        _lib.check(self.lib.fav_corrupt_normalize(
            self.handle.h, _ptr(x), _ptr(out), n, h, w, cfg.id, cfg.severity, int(seed), int(first_image),
            _lib.f3(self.mean), _lib.f3(self.std), flags, self._st()), "fav_corrupt_normalize")
        return out

    # ------------------------------------------------------------------ K2
    def forward_logits(self, x_bf16, T=1, p=0.2, seed=0, first_image=0, out=None):
        """bf16 [N,H,W,3] -> fp32 logits [N,T,C]."""
        n = x_bf16.shape[0]
        if out is None:
            out = torch.empty((n, T, self.num_classes), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.fav_forward_mc(self.handle.h, _ptr(x_bf16), _ptr(out), n, int(T), float(p), int(seed),
                                           int(first_image), self._st()), "fav_forward_mc")
        return out

    # ------------------------------------------------------------------ K3
    def epilogue(self, logits, labels=None, tau=0.9):
        n, T, c = logits.shape
        dev = self.device
        conf = torch.empty(n, dtype=torch.float32, device=dev)
        ent = torch.empty(n, dtype=torch.float32, device=dev)
        mi = torch.empty(n, dtype=torch.float32, device=dev)
        pred = torch.empty(n, dtype=torch.int32, device=dev)
        flag = torch.empty(n, dtype=torch.uint8, device=dev) if labels is not None else None
        _lib.check(self.lib.fav_epilogue(self.handle.h, _ptr(logits), _ptr(labels), n, T, c, float(tau), _ptr(conf),
                                         _ptr(ent), _ptr(mi), _ptr(pred), _ptr(flag), self._st()), "fav_epilogue")
        out = {"confidence": conf, "entropy": ent, "mutual_information": mi, "pred": pred}
        if flag is not None:
            out["failure_flag"] = flag
        return out

    # ------------------------------------------------------------------ public entry points
    def _labels(self, labels):
        if labels is None:
            return None
        if isinstance(labels, np.ndarray):
            labels = torch.from_numpy(labels)
        return labels.to(self.device, dtype=torch.int32).contiguous()

    def predict(self, images_u8, corruption=None, seed=0, first_image=0, bgr=False):
        """Deterministic (T=1, no dropout) prediction: {'logits','pred','confidence'}."""
        x = self.corrupt_normalize(images_u8, corruption, seed, first_image, bgr)
        logits = self.forward_logits(x, 1, 0.0, seed, first_image)
        u = self.epilogue(logits)
        return {"logits": logits[:, 0], "pred": u["pred"], "confidence": u["confidence"]}

    def uncertainty(self, images_u8, corruption=None, T=20, p=0.2, labels=None, tau=0.9, seed=0, first_image=0,
                    bgr=False):
        """MC-dropout uncertainty: {'confidence','entropy','mutual_information','pred'[, 'failure_flag']}."""
        x = self.corrupt_normalize(images_u8, corruption, seed, first_image, bgr)
        logits = self.forward_logits(x, T, p, seed, first_image)
        out = self.epilogue(logits, self._labels(labels), tau)
        out["logits"] = logits
        return out
