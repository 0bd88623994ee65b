"""GPU parity tests: every CUDA stage, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): corruption outputs within 1e-3 abs; logits
within bf16 tolerance; flags and integer histogram counts bit-exact apart from reported near-ties."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import corruptions as OC
from oracle import frame_stats as OFS
from oracle import metrics as OX
from oracle import model as OM
from oracle import philox as px
from oracle import sweep as OS
from oracle import uncertainty as OU

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_frames import frame_sequence  # noqa: E402


@pytest.fixture(scope="module")
def fav():
    import fav as _fav
    return _fav


@pytest.fixture(scope="module")
def clf18(fav):
    return fav.VisionClassifier("resnet18", 10, (32, 32), weights_seed=0, logit_gain=8.0)


@pytest.fixture(scope="module")
def folded18():
    return OM.fold_resnet(OM.build_torchvision("resnet18", 10, 0, logit_gain=8.0))


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------- RNG
def test_device_philox_streams_match_oracle(fav, clf18):
    for (n, h, w, first) in ((5, 32, 32, 0), (3, 17, 9, 7), (2, 224, 224, 123456)):
        d = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
        fav._lib.check(clf18.lib.fav_synth_images(clf18.handle.h, _p(d), n, h, w, 42, first, _s()), "synth")
        assert np.array_equal(d.cpu().numpy(), px.synthetic_images(n, h, w, 42, first))
    lab = torch.empty(1000, dtype=torch.int32, device="cuda")
    fav._lib.check(clf18.lib.fav_synth_labels(clf18.handle.h, _p(lab), 1000, 10, 9, 5, _s()), "labels")
    assert np.array_equal(lab.cpu().numpy(), px.synthetic_labels(1000, 10, 9, 5))


# ------------------------------------------------------------------------------------------- K1
ALL_CELLS = [(name, s) for name in OC.CORRUPTIONS for s in (1, 2, 3, 4, 5)]          # the whole 15 x 5 grid


def _k1_check(fav, clf, name, sev, x, seed, first, profile):
    """One cell: fp32 device output vs the oracle within 1e-3 abs (north_star bar); the two integer codecs byte-exact
    against PILLOW ITSELF (the third-party definition), not only against our restatement."""
    want = OC.corrupt(x, name, sev, seed=seed, first_image=first, profile=profile)
    cfg = fav.CorruptionConfig(name, sev)
    got = clf.corrupt_normalize(x, cfg, seed, first, out_f32=True, normalize=False).cpu().numpy()
    err = np.abs(got - want)
    assert err.max() <= 1e-3, f"{name} s{sev}: max abs err {err.max()}"
    c = OC.CONSTANTS[profile][name][sev - 1]
    if name == "jpeg_compression":
        from oracle import jpeg as OJ
        assert np.array_equal(np.rint(got * 255).astype(np.uint8), OJ.pil_roundtrip_u8(x, c)), "not byte-exact against PIL's JPEG"
        assert np.array_equal(got, want)
    if name == "pixelate":
        assert np.array_equal(np.rint(got * 255).astype(np.uint8), OC.pixelate_pil(x, c)), "not byte-exact against PIL's resize(BOX)"
        assert np.array_equal(got, want)
    return got, want


@pytest.mark.parametrize("name,sev", ALL_CELLS)
def test_k1_corruption_cifar_shape(fav, clf18, name, sev):
    n, first, seed = 12, 1000, 3
    x = px.synthetic_images(n, 32, 32, seed, first)
    got, want = _k1_check(fav, clf18, name, sev, x, seed, first, "cifar")
    # production output: bf16 of the normalised value (at most 1 bf16 ulp apart where fp32 rounding differs)
    gotb = clf18.corrupt_normalize(x, fav.CorruptionConfig(name, sev), seed, first).float().cpu().numpy()
    wantn = OC.normalize(want, *OC.MEAN_STD["cifar"])
    wantb = OC.to_bf16(wantn)
    assert (gotb == wantb).mean() > 0.99
    assert np.abs(gotb - wantn).max() <= 2.0 ** -7 * max(1.0, np.abs(wantn).max()) + 1e-3


@pytest.mark.parametrize("name,sev", ALL_CELLS)
def test_k1_corruption_imagenet_shape(fav, name, sev):
    """All 75 cells with the ImageNet-C constants at 224x224 (the wavefront swap chain at delta = 2 x 3 iterations and
    delta = 4, banded blurs at radius 6, elastic taps folded over the reflected row (r = 512), the largest defocus disk)."""
    clf = _clf_cache(fav, "resnet18", 1000, (224, 224))
    n, first, seed = 2, 77, 1
    x = px.synthetic_images(n, 224, 224, seed, first)
    _k1_check(fav, clf, name, sev, x, seed, first, "imagenet")


@pytest.mark.parametrize("hw", [(120, 160), (33, 47), (480, 640)])
def test_k1_ragged_frames_byte_exact_codecs(fav, hw):
    """Frames that are not a multiple of the 16x16 JPEG MCU (even and odd sizes: different bottom-edge rules in libjpeg)
    and the camera frame of config C5: jpeg_compression and pixelate stay byte-exact against Pillow."""
    clf = _clf_cache(fav, "resnet18", 1000, hw)
    x = px.synthetic_images(2, hw[0], hw[1], 5, 9)
    smooth = (np.add.outer(np.arange(hw[0]) * 3, np.arange(hw[1]) * 2)[..., None] + np.array([0, 40, 90])) % 256
    x[1] = smooth.astype(np.uint8)
    prof = OC.profile_for(*hw)
    for name in ("jpeg_compression", "pixelate"):
        for sev in (1, 3, 5):
            _k1_check(fav, clf, name, sev, x, 5, 9, prof)


@pytest.mark.parametrize("name,sev", ALL_CELLS)
def test_k1_corruption_camera_bgr_frames(fav, name, sev):
    """All 75 cells on 120x160 BGR frames (the layout the reference hands out, signal_analyzer.py:51,62; ImageNet-C constants;
    jpeg MCU rows that do not divide the frame): the device output in RGB order equals the oracle on the RGB image."""
    clf = _clf_cache(fav, "resnet18", 1000, (120, 160))
    n, first, seed = 2, 31, 8
    x = px.synthetic_images(n, 120, 160, seed, first)
    want = OC.corrupt(x, name, sev, seed=seed, first_image=first, profile="imagenet")
    got = clf.corrupt_normalize(np.ascontiguousarray(x[..., ::-1]), fav.CorruptionConfig(name, sev), seed, first, bgr=True,
                                out_f32=True, normalize=False).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-3, f"{name} s{sev}: max abs err {np.abs(got - want).max()}"


@pytest.mark.parametrize("bgr", [False, True])
@pytest.mark.parametrize("name", OC.CORRUPTIONS)
def test_k1_corruption_odd_small_frames(fav, name, bgr):
    """40x48 frames (CIFAR-10-C constants, 64-cell plasma maps in shared memory, 2x2 stencil tiles, a JPEG MCU row that is
    half padding), RGB and BGR sources, fp32 and production bf16 outputs."""
    clf = _clf_cache(fav, "resnet18", 10, (40, 48))
    n, first, seed, sev = 5, 3, 12, 3
    x = px.synthetic_images(n, 40, 48, seed, first)
    want = OC.corrupt(x, name, sev, seed=seed, first_image=first, profile="cifar")
    src = np.ascontiguousarray(x[..., ::-1]) if bgr else x
    cfg = fav.CorruptionConfig(name, sev)
    got = clf.corrupt_normalize(src, cfg, seed, first, bgr=bgr, out_f32=True, normalize=False).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-3, f"{name}: max abs err {np.abs(got - want).max()}"
    gotb = clf.corrupt_normalize(src, cfg, seed, first, bgr=bgr).float().cpu().numpy()
    wantn = OC.normalize(want, *OC.MEAN_STD["cifar"])
    assert (gotb == OC.to_bf16(wantn)).mean() > 0.99


def test_k1_fast_kernels_equal_legacy_kernels(fav, clf18):
    """The round-2 K1 kernels (chunk-per-thread clean / gaussian / impulse, guide-table shot noise, fused shared-memory plasma
    for fog / frost, raw-staged tap stencils with coalesced stores) produce the SAME BITS as the round-1 kernels they replace
    (fav_set_option 'k1_legacy')."""
    lib, h = clf18.lib, clf18.handle.h
    n, seed, first = 300, 7, 90
    x = torch.from_numpy(px.synthetic_images(n, 32, 32, seed, first)).cuda()
    for name in (None, "gaussian_noise", "impulse_noise", "contrast", "shot_noise", "fog", "frost", "defocus_blur", "motion_blur"):
        for sev in ((0,) if name is None else (1, 3, 5)):
            cfg = fav.CorruptionConfig(name, sev)
            fast = {f32: clf18.corrupt_normalize(x, cfg, seed, first, out_f32=f32, normalize=not f32) for f32 in (False, True)}
            fav._lib.check(lib.fav_set_option(h, b"k1_legacy", 1), "fav_set_option")
            try:
                for f32 in (False, True):
                    legacy = clf18.corrupt_normalize(x, cfg, seed, first, out_f32=f32, normalize=not f32)
                    assert torch.equal(fast[f32], legacy), (name, sev, f32)
            finally:
                fav._lib.check(lib.fav_set_option(h, b"k1_legacy", 0), "fav_set_option")


@pytest.mark.parametrize("hw,classes", [((224, 224), 1000), ((40, 48), 10), ((120, 160), 1000)])
def test_k1_dense_stencil_equals_list_stencil(fav, hw, classes):
    """defocus_blur's register-tiled dense loop (one staged pixel feeds four vertically adjacent outputs, weights of the tap
    box replicated per output row) adds the taps of every output in the list's row-major order: the SAME BITS as the
    tap-list loop ('k1_legacy'), on multi-tile frames, ragged tiles and the largest disk (radius 10, 21 x 21 box)."""
    clf = _clf_cache(fav, "resnet18", classes, hw)
    lib, h = clf.lib, clf.handle.h
    x = torch.from_numpy(px.synthetic_images(3, hw[0], hw[1], 4, 21)).cuda()
    for sev in (1, 2, 3, 4, 5):
        cfg = fav.CorruptionConfig("defocus_blur", sev)
        fast = {f32: clf.corrupt_normalize(x, cfg, 4, 21, out_f32=f32, normalize=not f32) for f32 in (False, True)}
        fav._lib.check(lib.fav_set_option(h, b"k1_legacy", 1), "fav_set_option")
        try:
            for f32 in (False, True):
                assert torch.equal(fast[f32], clf.corrupt_normalize(x, cfg, 4, 21, out_f32=f32, normalize=not f32)), (hw, sev, f32)
        finally:
            fav._lib.check(lib.fav_set_option(h, b"k1_legacy", 0), "fav_set_option")


def test_k1_table_taking_entry_equals_self_sufficient_entry(fav, clf18):
    """fav_corrupt_normalize_ex fed with fav_corrupt_params' host tables == fav_corrupt_normalize (which builds and caches the
    same tables inside the library), bit for bit, for every corruption."""
    lib, h = clf18.lib, clf18.handle.h
    n, seed, first = 6, 2, 40
    x = torch.from_numpy(px.synthetic_images(n, 32, 32, seed, first)).cuda()
    mean, std = fav._lib.f3(clf18.mean), fav._lib.f3(clf18.std)
    for name in OC.CORRUPTIONS:
        cid, sev = OC.CORRUPTION_ID[name], 4
        a = clf18.corrupt_normalize(x, fav.CorruptionConfig(name, sev), seed, first)
        fp, ip, tab = fav._lib.corrupt_params(cid, sev, 32, 32, "cifar")
        dtab = torch.from_numpy(tab).cuda() if tab is not None else None
        sb = int(lib.fav_corrupt_scratch_bytes(cid, n, 32, 32))
        scratch = torch.empty(max(sb, 1), dtype=torch.uint8, device="cuda")
        b = torch.empty_like(a)
        fa, ia = (C.c_float * max(1, len(fp)))(*fp), (C.c_int32 * max(1, len(ip)))(*ip)
        fav._lib.check(lib.fav_corrupt_normalize_ex(h, _p(x), _p(b), n, 32, 32, cid, sev, fa, len(fp), ia, len(ip), _p(dtab),
                                                    dtab.numel() if dtab is not None else 0, _p(scratch), sb, seed, first,
                                                    mean, std, 0, _s()), "fav_corrupt_normalize_ex")
        assert torch.equal(a, b), name


_CLFS = {}


def _clf_cache(fav, model, ncls, hw, gain=None):
    key = (model, ncls, hw, gain)
    if key not in _CLFS:
        _CLFS[key] = fav.VisionClassifier(model, ncls, hw, weights_seed=0, logit_gain=gain)
    return _CLFS[key]


def test_k1_bgr_and_ragged_and_clean(fav, clf18):
    x = px.synthetic_images(3, 32, 32, 5)
    a = clf18.corrupt_normalize(x, None, out_f32=True, normalize=False).cpu().numpy()
    assert np.array_equal(a, (x.astype(np.float32) / np.float32(255)))
    b = clf18.corrupt_normalize(np.ascontiguousarray(x[..., ::-1]), fav.CorruptionConfig("contrast", 3), 0, 0, bgr=True,
                                out_f32=True, normalize=False).cpu().numpy()
    assert np.abs(b - OC.corrupt(x, "contrast", 3)).max() <= 1e-6
    # empty batch is a no-op, bad severity raises
    clf18.corrupt_normalize(x[:0], fav.CorruptionConfig("gaussian_noise", 1))
    with pytest.raises(ValueError):
        fav.CorruptionConfig("gaussian_noise", 6)
    with pytest.raises(ValueError):
        fav.CorruptionConfig("no_such_corruption", 1)


def test_reset_releases_scratch_and_the_handle_keeps_working(fav):
    """fav_reset (signal_analyzer.py:41-45 style reset()): drops the forward workspace and the K1 scratch, keeps the weights and
    the per-cell K1 tables; the next calls re-grow what they need and give the same bits."""
    clf = fav.VisionClassifier("resnet18", 10, (32, 32), weights_seed=0, logit_gain=8.0)
    x = px.synthetic_images(9, 32, 32, 1)
    cfg = fav.CorruptionConfig("elastic_transform", 2)
    a = clf.uncertainty(x, cfg, T=3, labels=px.synthetic_labels(9, 10, 1), seed=1)
    clf.reset()
    b = clf.uncertainty(x, cfg, T=3, labels=px.synthetic_labels(9, 10, 1), seed=1)
    assert torch.equal(a["logits"], b["logits"]) and torch.equal(a["failure_flag"], b["failure_flag"])


def test_k1_partition_independence(fav, clf18):
    x = px.synthetic_images(10, 32, 32, 0)
    cfg = fav.CorruptionConfig("shot_noise", 2)
    whole = clf18.corrupt_normalize(x, cfg, 11, 0).cpu()
    part = clf18.corrupt_normalize(x[6:], cfg, 11, 6).cpu()
    assert torch.equal(whole[6:], part)


# ------------------------------------------------------------------------------------------- K2 conv
CONV_CASES = [
    # p, h, w, cin, cout, k, stride, pad, relu, res, modes
    (4, 8, 8, 64, 64, 3, 1, 1, 1, 1, (0, 1, 4)),
    (37, 8, 8, 64, 64, 3, 1, 1, 1, 0, (4,)),
    (3, 12, 5, 64, 64, 3, 1, 1, 0, 1, (0, 4)),
    (3, 8, 8, 64, 128, 3, 2, 1, 1, 0, (0, 1)),
    (3, 8, 8, 64, 128, 1, 2, 0, 0, 0, (0, 1)),
    (5, 28, 28, 128, 256, 3, 2, 1, 1, 1, (0, 1)),
    (2, 56, 56, 256, 512, 1, 2, 0, 0, 0, (0, 1)),
    (9, 2, 2, 256, 512, 3, 2, 1, 1, 0, (0, 1)),
    (17, 4, 4, 128, 128, 3, 1, 1, 1, 1, (0, 1)),
    (700, 4, 4, 128, 128, 3, 1, 1, 1, 1, (0,)),
    (40, 2, 2, 256, 256, 3, 1, 1, 1, 0, (0, 1)),
    (200, 1, 1, 512, 512, 3, 1, 1, 1, 1, (0, 1)),
    (5, 32, 32, 3, 64, 7, 2, 3, 1, 0, (2,)),
    (2, 14, 14, 256, 256, 3, 1, 1, 1, 1, (0, 1)),
    (2, 56, 56, 64, 64, 3, 1, 1, 0, 0, (0, 1)),
    (3, 7, 7, 512, 2048, 1, 1, 0, 0, 1, (0, 1)),
    (1, 9, 13, 16, 24, 3, 1, 1, 1, 0, (1, 2)),
    (1, 20, 160, 64, 64, 3, 1, 1, 1, 1, (0, 1)),          # wider than one A tile: rectangular pixel tiles
    (2, 30, 150, 64, 128, 3, 2, 1, 1, 0, (0, 1)),
    (3, 6, 6, 256, 512, 3, 1, 1, 1, 1, (0, 1)),           # 2-SM N = 256 tiles (auto)
    (1, 15, 20, 512, 512, 3, 1, 1, 1, 1, (0, 1)),         # few tiles, long K (split-K when FAV_SPLITK=1)
    (1, 30, 40, 256, 512, 3, 2, 1, 1, 0, (0,)),
    (2, 15, 20, 256, 100, 1, 1, 0, 0, 1, (0,)),
    (2, 56, 56, 64, 64, 3, 1, 1, 1, 1, (4,)),             # flat-padded kernel in band mode (image larger than a slab): bands of 2 rows
    (3, 57, 56, 64, 64, 3, 1, 1, 1, 0, (4,)),             # ... odd height: the last band is half empty
    (2, 30, 83, 64, 64, 3, 1, 1, 0, 1, (4,)),             # ... widest image that still fits: bands of 1 row
    (5, 28, 28, 64, 64, 3, 1, 1, 1, 1, (4,)),             # ... bands of 4 rows
    (11, 7, 7, 256, 512, 1, 1, 0, 1, 1, (0,)),            # flattened 1x1 on 7x7 images in the 2-SM kernel, odd number of pixel tiles
    (3, 14, 14, 128, 256, 1, 1, 0, 0, 1, (0,)),           # flattened 1x1, generic kernel
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_k2_conv_matches_torch_fp32(fav, clf18, case):
    p, h, w, cin, cout, k, stride, pad, relu, use_res, modes = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) % (2 ** 31))
    x = torch.randn((p, h, w, cin), generator=g).to(torch.bfloat16).cuda()
    wt = (torch.randn((cout, k, k, cin), generator=g) / (k * k * cin) ** 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    res = torch.randn((p, oh, ow, cout), generator=g).to(torch.bfloat16).cuda() if use_res else None
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, stride, pad)
    ref = ref.permute(0, 2, 3, 1)
    if res is not None:
        ref = ref + res.float()
    if relu:
        ref = torch.relu(ref)
    modes = tuple(modes) + ((0x200, 0x100) if 0 in modes else ())      # also force 256- and 128-pixel CTA tiles
    if 0 in modes and cout in (128, 512) and oh * ow <= 128:
        modes += (0x300,)                                               # 2-SM (cta_group::2) weights-stationary variant
    for mode in modes:
        for out_f32 in (0, 1):
            if mode == 4 and out_f32:
                continue          # the flat-padded variant only writes bf16 activations
            y = torch.full((p, oh, ow, cout), float("nan"), dtype=torch.float32 if out_f32 else torch.bfloat16, device="cuda")
            rc = clf18.lib.fav_conv2d(clf18.handle.h, _p(x), _p(wt), _p(bias), _p(res), _p(y), p, h, w, cin, cout, k, k,
                                      stride, pad, relu, out_f32, mode, _s())
            fav._lib.check(rc, "fav_conv2d")
            torch.cuda.synchronize()
            tol = 2e-3 if out_f32 else 2e-2
            err = (y.float() - ref).abs().max().item()
            assert not torch.isnan(y.float()).any(), f"mode {mode}: unwritten outputs"
            assert err <= tol * max(1.0, ref.abs().max().item()), f"mode {mode} f32={out_f32}: err {err}"


def test_k2_conv_splitk_matches_and_is_deterministic(fav):
    """Split-K (per-handle option, used by the batch-1 gate): same result as the unsplit kernel up to fp32 summation
    order, bit-identical from run to run (fixed-order fix-up by the last slice), tickets left clean for the next launch."""
    h2 = fav._lib.Handle(0)
    fav._lib.check(h2.lib.fav_set_option(h2.h, b"splitk", 1), "fav_set_option")
    assert h2.lib.fav_set_option(h2.h, b"no-such-option", 1) != 0
    for (p, h, w, cin, cout, k, stride, pad, relu, use_res) in ((1, 15, 20, 512, 512, 3, 1, 1, 1, 1), (1, 30, 40, 256, 512, 3, 2, 1, 1, 0),
                                                              (2, 15, 20, 256, 100, 1, 1, 0, 0, 1)):
        g = torch.Generator(device="cpu").manual_seed(p * 1000 + cin)
        x = torch.randn((p, h, w, cin), generator=g).to(torch.bfloat16).cuda()
        wt = (torch.randn((cout, k, k, cin), generator=g) / (k * k * cin) ** 0.5).to(torch.bfloat16).cuda()
        bias = torch.randn(cout, generator=g).cuda()
        oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        res = torch.randn((p, oh, ow, cout), generator=g).to(torch.bfloat16).cuda() if use_res else None
        ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, stride, pad).permute(0, 2, 3, 1)
        if res is not None:
            ref = ref + res.float()
        if relu:
            ref = torch.relu(ref)
        outs = []
        for _ in range(3):
            y = torch.full((p, oh, ow, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
            fav._lib.check(h2.lib.fav_conv2d(h2.h, _p(x), _p(wt), _p(bias), _p(res), _p(y), p, h, w, cin, cout, k, k, stride, pad,
                                             relu, 0, 0, _s()), "fav_conv2d")
            torch.cuda.synchronize()
            outs.append(y)
        assert not torch.isnan(outs[0].float()).any()
        assert (outs[0].float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


# ------------------------------------------------------------------------------------------- K2 forward
@pytest.mark.parametrize("T", [1, 4])
def test_k2_forward_logits_vs_oracle(fav, clf18, folded18, T):
    n, seed, first, p = 48, 2, 300, 0.25
    x = px.synthetic_images(n, 32, 32, seed, first)
    xc = OC.corrupt(x, "gaussian_noise", 2, seed=seed, first_image=first)
    xn = OC.to_bf16(OC.normalize(xc, *OC.MEAN_STD["cifar"]))
    xb = torch.from_numpy(xn).to(torch.bfloat16).cuda()
    got = clf18.forward_logits(xb, T, p, seed, first).cpu().numpy()
    assert got.shape == (n, T, 10)
    emu = OM.forward(folded18, xn, T=T, p=p, seed=seed, first_image=first, emulate_bf16=True)
    f32 = OM.forward(folded18, xn, T=T, p=p, seed=seed, first_image=first, emulate_bf16=False)
    scale = np.abs(f32).max()
    # same rounding points as the device: only fp32 accumulation order differs (rare bf16 rounding flips propagate)
    assert np.abs(got - emu).max() <= 2e-2 * scale, np.abs(got - emu).max() / scale
    # bf16 tolerance against the fp32 reference path
    assert np.abs(got - f32).max() <= 6e-2 * scale, np.abs(got - f32).max() / scale
    if T > 1:      # passes differ, and a different first_image gives different masks
        assert np.abs(got[:, 0] - got[:, 1]).max() > 1e-3
        other = clf18.forward_logits(xb, T, p, seed, first + 1).cpu().numpy()
        assert np.abs(other - got).max() > 1e-3
    # argmax agreement apart from near-ties (reported)
    pb_g, pb_r = OU.softmax(got).mean(1), OU.softmax(f32).mean(1)
    dis = pb_g.argmax(-1) != pb_r.argmax(-1)
    gap = OU.top2_gap(pb_r)
    assert (gap[dis] < 0.05).all(), f"argmax differs on non-tied samples: gaps {gap[dis]}"


@pytest.mark.parametrize("n,T", [(5, 3), (1, 2), (131, 1)])
def test_k2_forward_ragged_counts(fav, clf18, folded18, n, T):
    """Image counts that leave partial tiles everywhere (odd pass-image counts in the flat slabs, a lone image, a count
    that is not a multiple of any tile), against the bf16-emulating oracle."""
    seed, first, p = 3, 17, 0.2
    xn = OC.to_bf16(OC.normalize(px.synthetic_images(n, 32, 32, seed, first).astype(np.float32) / np.float32(255), *OC.MEAN_STD["cifar"]))
    got = clf18.forward_logits(torch.from_numpy(xn).to(torch.bfloat16).cuda(), T, p, seed, first).cpu().numpy()
    emu = OM.forward(folded18, xn, T=T, p=p, seed=seed, first_image=first, emulate_bf16=True)
    assert got.shape == (n, T, 10)
    assert np.abs(got - emu).max() <= 2e-2 * np.abs(emu).max(), np.abs(got - emu).max() / np.abs(emu).max()


def test_k2_forward_resnet50_small(fav):
    clf = _clf_cache(fav, "resnet50", 100, (64, 64), 4.0)
    folded = OM.fold_resnet(OM.build_torchvision("resnet50", 100, 0, logit_gain=4.0))
    n, T, p = 6, 3, 0.2
    xn = OC.to_bf16(np.random.default_rng(0).standard_normal((n, 64, 64, 3)).astype(np.float32))
    got = clf.forward_logits(torch.from_numpy(xn).to(torch.bfloat16).cuda(), T, p, 5, 0).cpu().numpy()
    emu = OM.forward(folded, xn, T=T, p=p, seed=5, emulate_bf16=True)
    assert np.abs(got - emu).max() <= 3e-2 * np.abs(emu).max()


# ------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n,T,Cc", [(257, 20, 10), (64, 1, 10), (33, 1, 1000), (20, 3, 1000), (50, 5, 100), (1, 30, 37)])
def test_k3_epilogue_vs_oracle(fav, clf18, n, T, Cc):
    rng = np.random.default_rng(n + T + Cc)
    z = (rng.standard_normal((n, T, Cc)) * 4).astype(np.float32)
    z[0, 0, :2] = 50.0            # an exact tie at the top -> lowest index wins
    y = rng.integers(0, Cc, n).astype(np.int32)
    out = clf18.epilogue(torch.from_numpy(z).cuda(), torch.from_numpy(y).cuda(), tau=0.3)
    ref = OU.uncertainty(z, y, 0.3)
    assert np.abs(out["confidence"].cpu().numpy() - ref["confidence"]).max() <= 2e-6
    assert np.abs(out["entropy"].cpu().numpy() - ref["entropy"]).max() <= 2e-5
    assert np.abs(out["mutual_information"].cpu().numpy() - ref["mutual_information"]).max() <= 2e-5
    pred = out["pred"].cpu().numpy()
    dis = pred != ref["pred"]
    assert (OU.top2_gap(ref["pbar"])[dis] < 1e-6).all()
    if T == 1:
        assert pred[0] == 0
    flag = out["failure_flag"].cpu().numpy()
    conf = out["confidence"].cpu().numpy()
    assert np.array_equal(flag, ((pred != y) & (conf >= np.float32(0.3))).astype(np.uint8))


# ------------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("n,Cc", [(10000, 10), (3000, 1000), (1, 10), (777, 100)])
def test_k4_histograms_bit_exact(fav, clf18, n, Cc):
    from fav.sweep import MetricsAccumulator, finalize
    clf = clf18 if Cc == 10 else _clf_cache(fav, "resnet18", Cc, (32, 32))
    rng = np.random.default_rng(n)
    conf = rng.uniform(0, 1, n).astype(np.float32)
    conf[: min(n, 16)] = (np.arange(min(n, 16)) / 15).astype(np.float32)         # exact bin edges
    H = rng.uniform(0, np.log(Cc), n).astype(np.float32)
    mi = rng.uniform(0, 1, n).astype(np.float32)
    y = rng.integers(0, Cc, n).astype(np.int32)
    pred = np.where(rng.uniform(size=n) < 0.6, y, rng.integers(0, Cc, n)).astype(np.int32)
    acc = MetricsAccumulator(clf, 2)
    t = lambda a: torch.from_numpy(a).cuda()
    acc.add_scores(1, t(conf), t(H), t(mi), t(pred), t(y), 0.7)
    acc.add_scores(1, t(conf[: n // 2]), t(H[: n // 2]), t(mi[: n // 2]), t(pred[: n // 2]), t(y[: n // 2]), 0.7)
    ar = np.zeros(OX.arena_words(Cc), np.int64)
    OX.accumulate(ar, conf, H, mi, pred, y, 0.7, Cc)
    OX.accumulate(ar, conf[: n // 2], H[: n // 2], mi[: n // 2], pred[: n // 2], y[: n // 2], 0.7, Cc)
    dev = acc.arena.cpu().numpy()
    assert np.array_equal(dev[1], ar)
    assert not dev[0].any()
    a, b = finalize(dev[1], Cc, 15, 4096), OX.finalize(ar, Cc)
    for k in b:
        assert a[k] == b[k] or (np.isnan(a[k]) and np.isnan(b[k]))


def test_k34_fused_equals_separate(fav, clf18):
    from fav.sweep import MetricsAccumulator
    rng = np.random.default_rng(3)
    z = torch.from_numpy((rng.standard_normal((5000, 7, 10)) * 3).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.integers(0, 10, 5000).astype(np.int32)).cuda()
    u = clf18.epilogue(z, y, 0.6)
    a = MetricsAccumulator(clf18, 1)
    a.add_scores(0, u["confidence"], u["entropy"], u["mutual_information"], u["pred"], y, 0.6)
    b = MetricsAccumulator(clf18, 1)
    outs = {k: torch.empty_like(v) for k, v in u.items()}
    b.add_logits(0, z, y, 0.6, outs)
    assert torch.equal(a.arena, b.arena)
    for k in u:
        assert torch.equal(u[k], outs[k])
    assert int(a.arena[0, 0]) == 5000


# ------------------------------------------------------------------------------------------- end to end
@pytest.mark.parametrize("name,sev,T", [("gaussian_noise", 3, 1), ("contrast", 2, 5), ("shot_noise", 4, 3)])
def test_cell_end_to_end_vs_oracle(fav, folded18, name, sev, T):
    from fav.sweep import CorruptionSweep, SweepConfig
    n, tau, p, seed = 96, 0.5, 0.2, 4
    x = px.synthetic_images(n, 32, 32, seed)
    y = px.synthetic_labels(n, 10, seed)
    cfg = SweepConfig(corruptions=(name,), severities=(sev,), T=T, p_drop=p, tau=tau, seed=seed, logit_gain=8.0, block=40)
    sw = CorruptionSweep(cfg)
    res = sw.run(x, y)[(name, sev)]
    u, ar = OS.eval_cell(folded18, x, y, name, sev, T=T, p=p, tau=tau, seed=seed, emulate_bf16=True)
    dev = sw.acc.arena[0].cpu().numpy()
    assert dev[0] == n == ar[0]
    # samples whose decision could legitimately differ: top-2 gap or distance to tau / bin edge within tolerance
    tol = 0.03
    gap = OU.top2_gap(u["pbar"])
    risky = int((gap < tol).sum())
    assert abs(int(dev[1]) - int(ar[1])) <= risky, (dev[1], ar[1], risky)
    near_tau = int((np.abs(u["confidence"] - tau) < tol).sum())
    assert abs(int(dev[2]) - int(ar[2])) <= risky + near_tau
    ref = OX.finalize(ar, 10)
    assert abs(res["mean_confidence"] - ref["mean_confidence"]) < 5e-3
    assert abs(res["mean_entropy"] - ref["mean_entropy"]) < 5e-3
    assert abs(res["ece"] - ref["ece"]) < 0.05
    print(f"[report] {name} s{sev} T={T}: near-tie samples={risky}, near-tau={near_tau}, "
          f"acc dev/oracle={dev[1]}/{ar[1]}, flags={dev[2]}/{ar[2]}, ece={res['ece']:.4f}/{ref['ece']:.4f}")


E2E_CELLS = [("gaussian_noise", 3), ("defocus_blur", 2), ("shot_noise", 2), ("contrast", 4), ("jpeg_compression", 3), ("fog", 5)]


def test_sweep_end_to_end_2048_images_vs_oracle(fav, folded18):
    """north_star: 'matching failure flags and ECE'.  2048 CIFAR-shape images x 6 cells (RNG, stencil, table, statistics,
    codec, two-pass) at T = 20 through the public sweep API, against the bf16-emulating oracle: |dECE| <= 0.01, AUROC x 3
    and mean MI within 0.01 / 2e-3, accuracy / flag counts equal apart from samples whose top-2 gap or distance to tau is
    inside the bf16 tolerance -- those samples are listed in the report (gpurun_out/parity_e2e_report.json and the log)."""
    from fav.sweep import CorruptionSweep, SweepConfig, HDR
    n, T, tau, p, seed = 2048, 20, 0.5, 0.2, 4
    x = px.synthetic_images(n, 32, 32, seed)
    y = px.synthetic_labels(n, 10, seed)
    cfg = SweepConfig(corruptions=tuple(dict.fromkeys(c for c, _ in E2E_CELLS)), severities=(1, 2, 3, 4, 5), T=T, p_drop=p,
                      tau=tau, seed=seed, logit_gain=8.0, block=512)
    sw = CorruptionSweep(cfg)
    sw.cells = [fav.CorruptionConfig(c, s) for c, s in E2E_CELLS]
    sw.acc = type(sw.acc)(sw.clf, len(sw.cells))
    res = sw.run(x, y)
    report = {"n": n, "T": T, "tau": tau, "cells": []}
    tol = 0.02                                    # bf16 tolerance on a pass-averaged probability
    checks = []
    for ci, (name, sev) in enumerate(E2E_CELLS):
        u, ar = OS.eval_cell(folded18, x, y, name, sev, T=T, p=p, tau=tau, seed=seed, emulate_bf16=True)
        ref = OX.finalize(ar, 10)
        r = res[(name, sev)]
        dev = sw.acc.arena[ci].cpu().numpy()
        d = sw.clf.uncertainty(x, fav.CorruptionConfig(name, sev), T=T, p=p, labels=y, tau=tau, seed=seed)
        dconf, dpred, dflag = d["confidence"].cpu().numpy(), d["pred"].cpu().numpy(), d["failure_flag"].cpu().numpy()
        gap = OU.top2_gap(u["pbar"])
        pred_diff = np.nonzero(dpred != u["pred"])[0]
        oflag = ((u["pred"] != y) & (u["confidence"] >= np.float32(tau))).astype(np.uint8)
        flag_diff = np.nonzero(dflag != oflag)[0]
        near_tau = np.abs(u["confidence"] - tau) < tol
        bins_d, bins_o = dev[HDR:HDR + 45].reshape(15, 3)[:, 0], ar[HDR:HDR + 45].reshape(15, 3)[:, 0]
        edge = np.abs(u["confidence"] * 15 - np.rint(u["confidence"] * 15)) < 15 * tol
        cell = {"cell": f"{name}/s{sev}", "near_tie_samples": [int(i) for i in pred_diff], "near_tie_gaps": [float(gap[i]) for i in pred_diff],
                "flag_diff_samples": [int(i) for i in flag_diff],
                "flag_diff_explained": [bool(gap[i] < 2 * tol or near_tau[i]) for i in flag_diff],
                "acc_dev_oracle": [int(dev[1]), int(ar[1])], "flags_dev_oracle": [int(dev[2]), int(ar[2])],
                "n_dev_oracle": [int(dev[0]), int(ar[0])], "ece_dev_oracle": [r["ece"], ref["ece"]],
                "auroc_dev_oracle": {k: [r[k], ref[k]] for k in ("auroc_msp", "auroc_entropy", "auroc_mi")},
                "mean_mi_dev_oracle": [r["mean_mutual_information"], ref["mean_mutual_information"]],
                "mean_entropy_dev_oracle": [r["mean_entropy"], ref["mean_entropy"]],
                "mean_conf_dev_oracle": [r["mean_confidence"], ref["mean_confidence"]],
                "max_abs_dconf": float(np.abs(dconf - u["confidence"]).max()),
                "max_abs_dH": float(np.abs(d["entropy"].cpu().numpy() - u["entropy"]).max()),
                "max_abs_dMI": float(np.abs(d["mutual_information"].cpu().numpy() - u["mutual_information"]).max()),
                "ece_bin_count_l1": int(np.abs(bins_d - bins_o).sum()), "samples_near_a_bin_edge": int(edge.sum())}
        report["cells"].append(cell)
        print("[parity report]", json.dumps(cell))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "parity_e2e_report.json"), "w") as fh:
        json.dump(report, fh, indent=1)
    for c in report["cells"]:                      # the report is on disk before anything can fail
        nm = c["cell"]
        assert c["n_dev_oracle"] == [n, n], nm
        assert c["max_abs_dconf"] <= tol and c["max_abs_dH"] <= 0.06 and c["max_abs_dMI"] <= 0.03, (nm, c["max_abs_dconf"], c["max_abs_dH"], c["max_abs_dMI"])
        assert all(g < 2 * tol for g in c["near_tie_gaps"]), f"{nm}: argmax differs on a non-tied sample, gaps {c['near_tie_gaps']}"
        assert all(c["flag_diff_explained"]), f"{nm}: failure flag differs away from any tie"
        # integer aggregates: equal apart from the reported samples
        assert abs(c["acc_dev_oracle"][0] - c["acc_dev_oracle"][1]) <= len(c["near_tie_samples"]), nm
        assert abs(c["flags_dev_oracle"][0] - c["flags_dev_oracle"][1]) <= len(c["flag_diff_samples"]), nm
        assert c["ece_bin_count_l1"] <= 2 * c["samples_near_a_bin_edge"], nm
        assert abs(c["ece_dev_oracle"][0] - c["ece_dev_oracle"][1]) <= 0.01, (nm, c["ece_dev_oracle"])
        assert abs(c["mean_mi_dev_oracle"][0] - c["mean_mi_dev_oracle"][1]) <= 2e-3, (nm, c["mean_mi_dev_oracle"])
        assert abs(c["mean_entropy_dev_oracle"][0] - c["mean_entropy_dev_oracle"][1]) <= 5e-3, nm
        assert abs(c["mean_conf_dev_oracle"][0] - c["mean_conf_dev_oracle"][1]) <= 5e-3, nm
        for k, (a_, b_) in c["auroc_dev_oracle"].items():
            assert abs(a_ - b_) <= 0.01, (nm, k, a_, b_)


def test_c_only_driver_matches_python_host(fav, tmp_path):
    """tests/c/c_abi_smoke.c drives K1 -> K2 -> K3+K4 from C alone (self-sufficient fav_corrupt_normalize: no Python tables);
    its per-cell arenas equal the Python host path's bit for bit (header fields + FNV-1a of the whole row)."""
    import subprocess
    from fav import weights
    from fav.sweep import CorruptionSweep, SweepConfig
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "bin", "c_abi_smoke")
    assert os.path.exists(exe), "tests/bin/c_abi_smoke is not built (failure-aware-vision_b200/csrc/build.sh)"
    blob = weights.pack_resnet(weights.build_model("resnet18", 10, 0, 8.0), "resnet18")
    wpath = tmp_path / "r18.favw"
    wpath.write_bytes(blob)
    N, T = 96, 4
    out = subprocess.run([exe, str(wpath), str(N), str(T)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("cell ")]
    assert len(rows) == 16
    sw = CorruptionSweep(SweepConfig(T=T, p_drop=0.2, tau=0.5, seed=6, logit_gain=8.0, block=N))
    sw.cells = [fav.CorruptionConfig(OC.CORRUPTIONS[int(r[1]) - 1] if int(r[1]) else None, int(r[2])) for r in rows]
    sw.acc = type(sw.acc)(sw.clf, len(sw.cells))
    sw.run(px.synthetic_images(N, 32, 32, 6), px.synthetic_labels(N, 10, 6))
    arena = sw.acc.arena.cpu().numpy()
    for i, r in enumerate(rows):
        f = dict(zip(r[3::2], r[4::2]))
        a = arena[i]
        assert [int(f[k]) for k in ("n", "correct", "flags", "sum_conf", "sum_h", "sum_mi")] == [int(v) for v in a[:6]], r
        fnv = 1469598103934665603
        for v in a.astype(np.uint64).tolist():
            fnv = ((fnv ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert int(f["fnv"]) == fnv, (r[1], r[2])


def test_sweep_cli_emits_the_wire_format(fav, tmp_path, capsys):
    """f3: `python -m fav.sweep --format json` (run in-process) writes one 'sweep_result' document whose per-cell records carry
    the metrics AND device time, evals/s, TFLOP/s and roofline fraction; the CSV form has the same columns."""
    from fav.sweep import main, CorruptionSweep
    out = tmp_path / "sweep.json"
    assert main(["--images", "96", "--passes", "3", "--block", "64", "--corruptions", "fog,jpeg_compression", "--severities", "2,5",
                 "--format", "json", "--out", str(out)]) == 0
    doc = json.loads(out.read_text())
    assert doc["type"] == "sweep_result" and doc["images"] == 96 and doc["n_gpus"] == 1 and doc["gflop_per_eval"] > 0
    assert [(c["corruption"], c["severity"]) for c in doc["cells"]] == [("fog", 2), ("fog", 5), ("jpeg_compression", 2), ("jpeg_compression", 5)]
    for c in doc["cells"]:
        assert c["n"] == 96 and 0 <= c["ece"] <= 1 and c["gpu_ms"] > 0 and c["evals_per_gpu_s"] > 0 and 0 < c["roofline_frac"] < 1
    assert main(["--images", "64", "--passes", "1", "--block", "64", "--corruptions", "contrast", "--severities", "3", "--format", "csv"]) == 0
    hdr = capsys.readouterr().out.strip().splitlines()[0].split(",")
    assert hdr == CorruptionSweep.COLUMNS + CorruptionSweep.PERF_COLUMNS


def _two_gpu_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import fav as _fav
    from fav.sweep import CorruptionSweep, SweepConfig
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = SweepConfig(corruptions=("impulse_noise", "brightness", "zoom_blur"), severities=(1, 4), T=5, logit_gain=8.0, block=64, seed=2)
    sw = CorruptionSweep(cfg, device=rank)
    sw.prepare(300)
    x, y = px.synthetic_images(300, 32, 32, 2), px.synthetic_labels(300, 10, 2)
    res = sw.run(x, y, rank=rank, world_size=world)               # the library's own ncclAllReduce (fav_allreduce) inside
    np.save(os.path.join(tmp, f"arena_r{rank}.npy"), sw.acc.arena.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_gpu_sweep_is_bit_identical_to_one_gpu(fav, tmp_path):
    """The real thing (not emulated ranks): two processes, two GPUs, work items dealt round-robin, one integer all-reduce
    over NCCL issued by the library -> every rank's arena equals the single-GPU arena bit for bit."""
    import torch.multiprocessing as mp
    from fav.sweep import CorruptionSweep, SweepConfig
    tmp = str(tmp_path)
    mp.start_processes(_two_gpu_worker, args=(2, 29641, tmp), nprocs=2, join=True, start_method="spawn")
    cfg = SweepConfig(corruptions=("impulse_noise", "brightness", "zoom_blur"), severities=(1, 4), T=5, logit_gain=8.0, block=64, seed=2)
    sw = CorruptionSweep(cfg)
    sw.run(px.synthetic_images(300, 32, 32, 2), px.synthetic_labels(300, 10, 2))
    one = sw.acc.arena.cpu().numpy()
    r0, r1 = np.load(os.path.join(tmp, "arena_r0.npy")), np.load(os.path.join(tmp, "arena_r1.npy"))
    assert np.array_equal(r0, r1) and np.array_equal(r0, one)
    assert one[:, 0].tolist() == [300] * 6


def test_sweep_partition_is_bit_identical(fav):
    """Emulated ranks on one GPU: the arenas of a 1-rank run and the sum of a 3-rank split are identical."""
    from fav.sweep import CorruptionSweep, SweepConfig, partition
    cfg = SweepConfig(corruptions=("impulse_noise", "brightness"), severities=(1, 4), T=3, logit_gain=8.0, block=32)
    x = px.synthetic_images(100, 32, 32, 0)
    y = px.synthetic_labels(100, 10, 0)
    sw = CorruptionSweep(cfg)
    sw.run(x, y)
    whole = sw.acc.arena.clone()
    total = torch.zeros_like(whole)
    for r in range(3):
        sw.reset()
        xd, yd = sw.clf._images(x), sw.clf._labels(y)
        items = sw.work_items(100)
        for i in partition(len(items), r, 3):
            sw.run_item(xd, yd, items[i])
        total += sw.acc.arena
    assert torch.equal(total, whole)
    assert int(whole[:, 0].sum()) == 100 * 4
    # the host-streaming API reports every step's arena row (one step late); the last row of a cell is its final state
    sw.reset()
    rows = {}
    hx, hy = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).to(torch.int32).pin_memory()
    n = sw.run_stream(hx, hy, sw.work_items(100), on_row=lambda item, row: rows.__setitem__(item[0], row.clone()))
    assert n == 100 * 4 and torch.equal(sw.acc.arena, whole)
    for ci, row in rows.items():
        assert torch.equal(row, whole[ci].cpu())
    # the pipelined resident path (K1 of step k+1 on a side stream beside the forward of step k) and the step-by-step path
    sw.reset()
    assert sw.run_items(xd, yd, sw.work_items(100)) == 100 * 4 and torch.equal(sw.acc.arena, whole)
    sw.reset()
    for it in sw.work_items(100):
        sw.run_item(xd, yd, it)
    assert torch.equal(sw.acc.arena, whole)


def test_allreduce_entry_points_single_rank(fav, clf18):
    """fav_allreduce is a no-op on one rank (the N-rank exchange is checked by tools/allreduce_check.py under torchrun and
    by the multi-GPU bench); the unique-id entry point reaches NCCL."""
    a = torch.arange(1000, dtype=torch.int64, device="cuda")
    fav._lib.check(clf18.lib.fav_allreduce(clf18.handle.h, _p(a), a.numel(), _s()), "fav_allreduce")
    torch.cuda.synchronize()
    assert torch.equal(a.cpu(), torch.arange(1000, dtype=torch.int64))
    buf = C.create_string_buffer(128)
    fav._lib.check(clf18.lib.fav_comm_unique_id(buf), "fav_comm_unique_id")
    assert any(buf.raw)
    fav._lib.check(clf18.lib.fav_comm_init(clf18.handle.h, buf, 0, 1), "fav_comm_init")       # world of one: no communicator


# ------------------------------------------------------------------------------------------- f1 + gate
def test_frame_stats_kernel_bit_exact(fav, clf18):
    for seed, (h, w) in ((0, (240, 320)), (1, (480, 640)), (2, (96, 130)), (3, (2, 2)), (4, (33, 1031))):
        frames = frame_sequence(seed, h, w)[:4] if h > 2 else [np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(3)]
        gray_d = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        out = torch.zeros(260, dtype=torch.int64, device="cuda")
        prev = None
        for i, f in enumerate(frames):
            fd = torch.from_numpy(f).cuda()
            fav._lib.check(clf18.lib.fav_frame_stats(clf18.handle.h, _p(fd), _p(gray_d), h, w, 1 if i == 0 else 0, _p(out), _s()), "stats")
            want, gray = OFS.frame_stats(f, prev)
            prev = gray
            assert np.array_equal(out.cpu().numpy(), want), (seed, i)
            assert np.array_equal(gray_d.cpu().numpy(), gray)


def test_gate_reproduces_reference_signal_analyzer(fav):
    with open(os.path.join(os.path.dirname(__file__), "golden", "signal_analyzer.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"][:2]:
        gate = fav.UncertaintyGate(frame_hw=(case["h"], case["w"]), score_source="signal", use_classifier=False)
        for f, want in zip(frame_sequence(case["seed"], case["h"], case["w"]), case["results"]):
            got = gate.analyze_frame(f)
            assert got == want


def test_gate_with_classifier_feeds_trust_engine_contract(fav):
    gate = fav.UncertaintyGate(frame_hw=(96, 128), T=4, num_classes=10, logit_gain=8.0)
    f = frame_sequence(0, 96, 128)
    r = gate.analyze_frame(f[1])
    assert set(r) == {"anomaly_score", "vision_status", "metrics"}
    assert set(r["metrics"]) >= {"blur", "brightness", "freeze", "entropy", "raw", "uncertainty"}
    assert 0.0 <= r["anomaly_score"] <= 1.0 and r["vision_status"].startswith("VISION_")
    u = r["metrics"]["uncertainty"]
    assert 0 < u["confidence"] <= 1 and u["entropy"] >= 0 and u["mutual_information"] >= 0
    gate.reset()
    assert gate.analyze_frame(f[1])["metrics"]["raw"]["frame_diff"] == 10.0


# ------------------------------------------------------------------------------------------- BASELINE shapes C3 / C4 / C5
@pytest.mark.parametrize("T", [1, 3])
def test_k2_forward_resnet50_imagenet_shape(fav, T):
    """ResNet-50 at 224x224, 1000 classes (configs C3 / C4): space-to-depth TMA stem, rectangular pixel tiles, staged
    epilogue, identity-MMA residuals and the 2-SM tiles at their real shapes, against the bf16-emulating oracle."""
    clf = _clf_cache(fav, "resnet50", 1000, (224, 224), 4.0)
    folded = OM.fold_resnet(OM.build_torchvision("resnet50", 1000, 0, logit_gain=4.0))
    n, p = 3, 0.2
    xn = OC.to_bf16(np.random.default_rng(1).standard_normal((n, 224, 224, 3)).astype(np.float32))
    got = clf.forward_logits(torch.from_numpy(xn).to(torch.bfloat16).cuda(), T, p, 9, 40).cpu().numpy()
    emu = OM.forward(folded, xn, T=T, p=p, seed=9, first_image=40, emulate_bf16=True)
    assert got.shape == (n, T, 1000)
    assert np.abs(got - emu).max() <= 3e-2 * np.abs(emu).max(), np.abs(got - emu).max() / np.abs(emu).max()


def test_k2_forward_resnet18_camera_shape(fav):
    """ResNet-18 on a 120x160 frame (the C5 gate's geometry at quarter size: feature maps wider than one 128-pixel tile)."""
    clf = _clf_cache(fav, "resnet18", 1000, (120, 160), 2.0)
    folded = OM.fold_resnet(OM.build_torchvision("resnet18", 1000, 0, logit_gain=2.0))
    xn = OC.to_bf16(np.random.default_rng(2).standard_normal((2, 120, 160, 3)).astype(np.float32))
    got = clf.forward_logits(torch.from_numpy(xn).to(torch.bfloat16).cuda(), 1, 0.0, 0, 0).cpu().numpy()
    emu = OM.forward(folded, xn, T=1, emulate_bf16=True)
    assert np.abs(got - emu).max() <= 3e-2 * np.abs(emu).max(), np.abs(got - emu).max() / np.abs(emu).max()


@pytest.mark.parametrize("T", [1, 20])
def test_k2_forward_resnet18_c5_geometry(fav, T):
    """Config C5 at its real geometry: ResNet-18 on one 480x640 frame, T = 1 (the graph-replayed gate) and T = 20 (MC-dropout
    gate), against the bf16-emulating oracle; plus K1 on the same BGR frame."""
    clf = _clf_cache(fav, "resnet18", 1000, (480, 640), 2.0)
    folded = OM.fold_resnet(OM.build_torchvision("resnet18", 1000, 0, logit_gain=2.0))
    frame = px.synthetic_images(1, 480, 640, 21, 5)                      # RGB
    bgr = np.ascontiguousarray(frame[..., ::-1])
    xk = clf.corrupt_normalize(bgr, None, 0, 0, bgr=True).float().cpu().numpy()
    xn = OC.to_bf16(OC.normalize(frame.astype(np.float32) / np.float32(255), *OC.MEAN_STD["imagenet"]))
    assert np.array_equal(xk, xn)                                        # clean K1 + BGR swap is exact
    got = clf.forward_logits(torch.from_numpy(xn).to(torch.bfloat16).cuda(), T, 0.2, 3, 11).cpu().numpy()
    emu = OM.forward(folded, xn, T=T, p=0.2, seed=3, first_image=11, emulate_bf16=True)
    assert got.shape == (1, T, 1000)
    assert np.abs(got - emu).max() <= 3e-2 * np.abs(emu).max(), np.abs(got - emu).max() / np.abs(emu).max()
    if T > 1:
        assert np.abs(got[:, 0] - got[:, 1]).max() > 1e-3               # passes differ


def test_k2_forward_resnet50_c4_shape_t30(fav):
    """Config C4 at its real pass count: ResNet-50, 224x224, T = 30 MC-dropout passes of one image, against the bf16-emulating
    oracle (the T masked replicas of block 0, dropout epilogues in every residual block, 1000-class fc)."""
    clf = _clf_cache(fav, "resnet50", 1000, (224, 224), 4.0)
    folded = OM.fold_resnet(OM.build_torchvision("resnet50", 1000, 0, logit_gain=4.0))
    xn = OC.to_bf16(np.random.default_rng(4).standard_normal((1, 224, 224, 3)).astype(np.float32))
    got = clf.forward_logits(torch.from_numpy(xn).to(torch.bfloat16).cuda(), 30, 0.2, 6, 123).cpu().numpy()
    emu = OM.forward(folded, xn, T=30, p=0.2, seed=6, first_image=123, emulate_bf16=True)
    assert got.shape == (1, 30, 1000)
    assert np.abs(got - emu).max() <= 3e-2 * np.abs(emu).max(), np.abs(got - emu).max() / np.abs(emu).max()
    assert np.abs(got[0, 0] - got[0, 29]).max() > 1e-3


def test_cell_end_to_end_resnet50_imagenet_shape(fav):
    """One C3-shaped cell end to end (ResNet-50, 224x224, 1000 classes, T = 1, 48 images, defocus_blur s3 with the ImageNet-C
    constants) through the sweep API against the oracle: per-class aggregate layout, ECE / confidence / entropy agreement,
    accuracy and flags equal apart from near-ties."""
    from fav.sweep import CorruptionSweep, SweepConfig
    n, tau, seed = 48, 0.5, 9
    x = px.synthetic_images(n, 224, 224, seed)
    y = px.synthetic_labels(n, 1000, seed)
    cfg = SweepConfig(model="resnet50", num_classes=1000, input_hw=(224, 224), corruptions=("defocus_blur",), severities=(3,), T=1,
                      tau=tau, seed=seed, logit_gain=4.0, block=32)
    sw = CorruptionSweep(cfg, classifier=_clf_cache(fav, "resnet50", 1000, (224, 224), 4.0))
    res = sw.run(x, y)[("defocus_blur", 3)]
    folded = OM.fold_resnet(OM.build_torchvision("resnet50", 1000, 0, logit_gain=4.0))
    u, ar = OS.eval_cell(folded, x, y, "defocus_blur", 3, T=1, tau=tau, seed=seed, num_classes=1000, emulate_bf16=True)
    dev = sw.acc.arena[0].cpu().numpy()
    ref = OX.finalize(ar, 1000)
    risky = int((OU.top2_gap(u["pbar"]) < 0.03).sum())
    assert dev[0] == ar[0] == n and abs(int(dev[1]) - int(ar[1])) <= risky
    assert abs(int(dev[2]) - int(ar[2])) <= risky + int((np.abs(u["confidence"] - tau) < 0.03).sum())
    assert abs(res["mean_confidence"] - ref["mean_confidence"]) < 5e-3 and abs(res["mean_entropy"] - ref["mean_entropy"]) < 2e-2
    assert abs(res["ece"] - ref["ece"]) < 0.02 and res["mean_mutual_information"] == 0.0
    per_class = dev[8 + 45 + 6 * 4096:].reshape(1000, 2)
    assert per_class[:, 0].sum() == n and np.array_equal(per_class[:, 0], np.bincount(y, minlength=1000))


def test_gate_graph_replay_equals_eager_launches(fav):
    """The CUDA-graph replay of the per-frame device work returns exactly what the eager launches return: T = 1, and T = 4 with
    frozen masks (the graph freezes first_image, the frozen schedule keeps it at 0 in eager mode too)."""
    frames = frame_sequence(3, 96, 128)[:8]
    for T, sched in ((1, "per_frame"), (4, "frozen")):
        clf = _clf_cache(fav, "resnet18", 10, (96, 128), 8.0)
        g = fav.UncertaintyGate(classifier=clf, frame_hw=(96, 128), T=T, mask_schedule=sched, use_graph=True)
        e = fav.UncertaintyGate(classifier=clf, frame_hw=(96, 128), T=T, mask_schedule=sched, use_graph=False)
        outs_g = [g.analyze_frame(f) for f in frames]
        assert g.graph_active and g.graph_error is None
        outs_e = [e.analyze_frame(f) for f in frames]
        assert not e.graph_active
        if T == 1 or sched == "frozen":
            assert outs_g == outs_e


# ------------------------------------------------------------------------------------------- f4: trust replay
def test_trust_replay_kernel_reproduces_the_reference_engine(fav):
    """The CUDA replay against trajectories of the REAL reference TrustEngine (golden, pinned): float64 state bit for
    bit, policy / contradiction outputs exactly; plus the oracle on a larger random batch and edge shapes."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_trust_inputs import sequences
    from oracle import trust as OT
    tr = fav.TrustReplay()
    with open(os.path.join(os.path.dirname(__file__), "golden", "trust_replay.json")) as fh:
        gold = json.load(fh)
    for c in gold["cases"]:
        status, score = sequences(c["seed"], c["n_seq"], c["n_ticks"])
        res = tr.run(status, score, c["dt"])
        t = np.array(c["trajectories"], dtype=np.float64)
        assert np.array_equal(res["state"], t[:, :, :5])
        assert np.array_equal(res["policy"], t[:, :, 5]) and np.array_equal(res["contradiction"], t[:, :, 6])
        assert np.array_equal(res["contradiction_count"], t[:, :, 7])
        assert np.array_equal(res["final"][:, :5], t[:, -1, :5]) and np.array_equal(res["final"][:, 5:], t[:, -1, 5:8])
        d = tr.state_dict(res, 0, c["n_ticks"] - 1, "VISION_OK")
        assert d["reliability"] == t[0, -1, 8] and d["trust_velocity"] == t[0, -1, 9] and d["tick_count"] == c["n_ticks"]
    status, score = sequences(7, 300, 257)                       # more sequences than one CTA, ragged against the block size
    dts = np.random.default_rng(7).uniform(0.005, 0.1, 257)
    a, b = tr.run(status, score, dts), OT.replay(status, score, dts)
    for k in ("state", "policy", "contradiction", "contradiction_count"):
        assert np.array_equal(a[k], b[k]), k
    only_final = tr.run(status, score, dts, trajectory=False)
    assert set(only_final) == {"final"} and np.array_equal(only_final["final"], a["final"])
    assert tr.run(np.zeros((0, 5), np.int8), np.zeros((0, 5)), 0.1)["final"].shape == (0, 8)


# ------------------------------------------------------------------------------------------- full BASELINE sizes
def test_c2_full_size_properties(fav):
    """Config C2 at its full size (N = 10 000 CIFAR-shape images, T = 20) on a few cells: size-independent properties of the
    aggregates -- every histogram family sums to N, integer sums are consistent, and the result does not depend on the block
    size used to walk the images (Philox counters are keyed by the global image index)."""
    from fav.sweep import CorruptionSweep, SweepConfig, HDR
    N, T = 10_000, 20
    cells = (("gaussian_noise",), (2, 5))
    arenas = []
    for block in (512, 384):
        cfg = SweepConfig(corruptions=cells[0], severities=cells[1], T=T, logit_gain=8.0, block=block, seed=11)
        sw = CorruptionSweep(cfg)
        h, st = sw.clf.handle.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)
        x = torch.empty((N, 32, 32, 3), dtype=torch.uint8, device="cuda")
        y = torch.empty(N, dtype=torch.int32, device="cuda")
        fav._lib.check(sw.clf.lib.fav_synth_images(h, _p(x), N, 32, 32, 11, 0, st), "synth")
        fav._lib.check(sw.clf.lib.fav_synth_labels(h, _p(y), N, 10, 11, 0, st), "synth")
        res = sw.run(x, y)
        arenas.append(sw.acc.arena.cpu().numpy())
        for (name, sev), r in res.items():
            assert r["n"] == N and 0 <= r["ece"] <= 1 and 0 <= r["accuracy"] <= 1
    a = arenas[0]
    assert np.array_equal(a, arenas[1]), "aggregates depend on the block size"
    for row in a:
        assert row[0] == N
        bins = row[HDR:HDR + 45].reshape(15, 3)
        assert bins[:, 0].sum() == N and bins[:, 2].sum() == row[1] and bins[:, 1].sum() == row[3]
        buckets = row[HDR + 45:HDR + 45 + 6 * 4096].reshape(3, 4096, 2)
        assert (buckets.sum(axis=(1, 2)) == N).all() and (buckets[:, :, 0].sum(1) == row[1]).all()
        conf = row[HDR + 45 + 6 * 4096:].reshape(10, 10)
        assert conf.sum() == N and np.trace(conf) == row[1]
        assert row[2] <= N - row[1]                       # flags are a subset of the misclassified samples


@pytest.mark.parametrize("cfgname,N,T,blocks", [("C3", 8192, 1, (256, 192)), ("C4", 1536, 30, (64, 48))])
def test_c3_c4_smoke_size_properties(fav, cfgname, N, T, blocks):
    """Configs C3 (ResNet-50 224x224, T = 1) and C4 (T = 30, mutual information + AUROC) at their smoke sizes:
    histogram families sum to N, per-class totals are consistent, MI is exactly 0 for T = 1, and the aggregates do not depend
    on the block size (images are generated on the device from the Philox stream; 1000 classes -> per-class confusion)."""
    from fav.sweep import CorruptionSweep, SweepConfig, HDR
    arenas = []
    for block in blocks:
        cfg = SweepConfig(model="resnet50", num_classes=1000, input_hw=(224, 224), corruptions=("brightness",), severities=(3,),
                          T=T, logit_gain=4.0, block=block, seed=5)
        sw = CorruptionSweep(cfg, classifier=_clf_cache(fav, "resnet50", 1000, (224, 224), 4.0))
        h, st = sw.clf.handle.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)
        x = torch.empty((N, 224, 224, 3), dtype=torch.uint8, device="cuda")
        y = torch.empty(N, dtype=torch.int32, device="cuda")
        fav._lib.check(sw.clf.lib.fav_synth_images(h, _p(x), N, 224, 224, 5, 0, st), "synth")
        fav._lib.check(sw.clf.lib.fav_synth_labels(h, _p(y), N, 1000, 5, 0, st), "synth")
        res = sw.run(x, y)
        arenas.append(sw.acc.arena.cpu().numpy())
        r = res[("brightness", 3)]
        assert r["n"] == N and 0 <= r["ece"] <= 1 and 0 <= r["auroc_msp"] <= 1
        if T == 1:
            assert r["mean_mutual_information"] == 0.0
        del x
    a = arenas[0]
    assert np.array_equal(a, arenas[1]), f"{cfgname}: aggregates depend on the block size"
    row = a[0]
    assert row[0] == N
    bins = row[HDR:HDR + 45].reshape(15, 3)
    assert bins[:, 0].sum() == N and bins[:, 2].sum() == row[1]
    buckets = row[HDR + 45:HDR + 45 + 6 * 4096].reshape(3, 4096, 2)
    assert (buckets.sum(axis=(1, 2)) == N).all() and (buckets[:, :, 0].sum(1) == row[1]).all()
    per_class = row[HDR + 45 + 6 * 4096:].reshape(1000, 2)
    assert per_class[:, 0].sum() == N and per_class[:, 1].sum() == row[1] and (per_class[:, 1] <= per_class[:, 0]).all()
