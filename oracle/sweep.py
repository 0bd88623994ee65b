"""End-to-end corruption-sweep cell on the CPU -- oracle.  TEST INFRASTRUCTURE ONLY.

This *is* "the reference PyTorch path" for parity and for the CPU timing baseline
(BASELINE.md section 3): the reference itself has no such code.
Stages: corrupt -> normalize -> ResNet x T -> uncertainty -> failure flag -> aggregates.
"""
import numpy as np

from . import corruptions as C
from . import metrics as X
from . import model as M
from . import uncertainty as U


def eval_cell(folded, x_u8, labels, corruption, severity, *, T=1, p=0.2, tau=0.9, seed=0,
              first_image=0, num_classes=10, mean_std=None, emulate_bf16=False, arena=None,
              profile=None):
    """Returns (per-sample dict, arena int64).  x_u8 [N,H,W,3] RGB."""
    n, h, w, _ = x_u8.shape
    mean, std = mean_std or C.MEAN_STD[C.profile_for(h, w)]
    xc = C.corrupt(x_u8, corruption, severity, seed=seed, first_image=first_image, profile=profile)
    xn = C.normalize(xc, mean, std)
    logits = M.forward(folded, xn, T=T, p=p, seed=seed, first_image=first_image,
                       emulate_bf16=emulate_bf16)
    u = U.uncertainty(logits, labels, tau)
    if arena is None:
        arena = np.zeros(X.arena_words(num_classes), dtype=np.int64)
    X.accumulate(arena, u["confidence"], u["entropy"], u["mutual_information"], u["pred"],
                 labels, tau, num_classes)
    u["logits"] = logits
    return u, arena
