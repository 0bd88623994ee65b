"""Corruption / severity configuration.

Reference surface being extended: ``VisionSimulator.set_noise / set_brightness / set_mode``
(platform/backend/vision_simulator.py:25-36) -- two sliders and four modes.  Here the config is
the 15-corruption x 5-severity grid of Hendrycks & Dietterich.

The per-cell kernel constants and tables (Poisson thresholds, stencil taps, resampling ranges,
libjpeg / Pillow coefficients ...) are built INSIDE the library (csrc/tables.cu), so the C ABI is
callable without this module; ``SEVERITY`` below is the human-readable copy of the same grid
(tests/test_host.py checks it against ``fav_corruption_constants``).  Product code: it never
imports ``oracle``.
"""
from dataclasses import dataclass

CORRUPTIONS = (
    "gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur",
    "motion_blur", "zoom_blur", "snow", "frost", "fog", "brightness", "contrast",
    "elastic_transform", "pixelate", "jpeg_compression",
)
CORRUPTION_ID = {name: i + 1 for i, name in enumerate(CORRUPTIONS)}
# corruptions with a device kernel in this build; the sweep reports the rest as unavailable
IMPLEMENTED = CORRUPTIONS      # all 15 have device kernels

SEVERITY = {
    "imagenet": {
        "gaussian_noise": [.08, .12, .18, .26, .38],
        "shot_noise": [60, 25, 12, 5, 3],
        "impulse_noise": [.03, .06, .09, .17, .27],
        "defocus_blur": [(3, .1), (4, .5), (6, .5), (8, .5), (10, .5)],
        "motion_blur": [(10, 3), (15, 5), (15, 8), (15, 12), (20, 15)],
        "zoom_blur": [(1.11, .01), (1.16, .01), (1.21, .02), (1.26, .02), (1.33, .03)],
        "fog": [(1.5, 2), (2., 2), (2.5, 1.7), (2.5, 1.5), (3., 1.4)],
        "brightness": [.1, .2, .3, .4, .5],
        "contrast": [.4, .3, .2, .1, .05],
        "pixelate": [.6, .5, .4, .3, .25],
        "jpeg_compression": [25, 18, 15, 10, 7],
        "glass_blur": [(.7, 1, 2), (.9, 2, 1), (1, 2, 3), (1.1, 3, 2), (1.5, 4, 2)],
        "frost": [(1, .4), (.8, .6), (.7, .7), (.65, .7), (.6, .75)],
        "snow": [(.1, .3, 3, .5, 10, 4, .8), (.2, .3, 2, .5, 12, 4, .7), (.55, .3, 4, .9, 12, 8, .7),
                 (.55, .3, 4.5, .85, 12, 8, .65), (.55, .3, 2.5, .85, 12, 12, .55)],
        "elastic_transform": [(2., .7, .1), (2., .08, .2), (.05, .01, .02), (.07, .01, .02), (.12, .01, .02)],
    },
    "cifar": {
        "gaussian_noise": [.04, .06, .08, .09, .10],
        "shot_noise": [500, 250, 100, 75, 50],
        "impulse_noise": [.01, .02, .03, .05, .07],
        "defocus_blur": [(.3, .4), (.4, .5), (.5, .6), (1, .2), (1.5, .1)],
        "motion_blur": [(10, 1), (10, 1.5), (10, 2), (10, 2.5), (12, 3)],
        "zoom_blur": [(1.06, .01), (1.11, .01), (1.16, .01), (1.21, .01), (1.26, .01)],
        "fog": [(.2, 3), (.5, 3), (.75, 2.5), (1, 2), (1.5, 1.75)],
        "brightness": [.05, .1, .15, .2, .3],
        "contrast": [.75, .5, .4, .3, .15],
        "pixelate": [.95, .9, .85, .75, .65],
        "jpeg_compression": [80, 65, 58, 50, 40],
        "glass_blur": [(.05, 1, 1), (.25, 1, 1), (.4, 1, 1), (.25, 1, 2), (.4, 1, 2)],
        "frost": [(1, .2), (1, .3), (.9, .4), (.85, .4), (.75, .45)],
        "snow": [(.1, .2, 1, .6, 8, 3, .95), (.1, .2, 1, .5, 10, 4, .9), (.15, .3, 1.75, .55, 10, 4, .9),
                 (.25, .3, 2.25, .6, 12, 6, .85), (.3, .3, 1.25, .65, 14, 12, .8)],
        "elastic_transform": [(0, 0, .08), (.05, .2, .07), (.08, .06, .06), (.1, .04, .05), (.1, .03, .03)],
    },
}

MEAN_STD = {
    "imagenet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "cifar": ((0.4914, 0.4822, 0.4465), (0.2470, 0.2435, 0.2616)),
}


def profile_for(h, w):
    """CIFAR-10-C constants for frames up to 64 px, ImageNet-C constants above."""
    return "cifar" if max(h, w) <= 64 else "imagenet"


@dataclass(frozen=True)
class CorruptionConfig:
    """One cell of the sweep grid.  ``name`` None / 'clean' with severity 0 is the clean cell.

    Mirrors the reference's permissive style (vision_simulator.py:27 silently ignores unknown
    modes) only in spirit: here an unknown name or severity raises, because a silently-ignored
    cell would corrupt a benchmark table."""
    name: str = None
    severity: int = 0

    def __post_init__(self):
        if self.name in (None, "clean", "none"):
            object.__setattr__(self, "name", None)
            object.__setattr__(self, "severity", 0)
            return
        if self.name not in CORRUPTION_ID:
            raise ValueError(f"unknown corruption '{self.name}'; expected one of {CORRUPTIONS}")
        if not 1 <= int(self.severity) <= 5:
            raise ValueError("severity must be in 1..5")
        object.__setattr__(self, "severity", int(self.severity))

    @property
    def id(self):
        return 0 if self.name is None else CORRUPTION_ID[self.name]

    def to_dict(self):
        return {"corruption": self.name or "clean", "severity": self.severity}
