"""Pins the oracle itself (CPU only): Philox KATs, ResNet restatement vs torchvision, AUROC vs
sklearn, ECE definition, corruption invariants.  The reference has no tests for this path
(SURVEY.md section 4), so these are the anchors that exist."""
import os

import numpy as np
import pytest

from oracle import corruptions as C
from oracle import metrics as X
from oracle import model as M
from oracle import philox as px
from oracle import uncertainty as U


def test_philox_random123_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kats = [
        ((0, 0, 0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0x299f31d0 << 32) | 0xa4093822,
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        got = tuple(int(v) for v in px.philox4x32_10(*ctr, key))
        assert got == want


def test_uniform_and_normal_moments():
    x = px.philox4x32_10(np.arange(50000), 7, 0, 99, 1234)
    u = px.u32_to_uniform(x[0])
    assert 0 < u.min() and u.max() <= 1.0 and abs(u.mean() - 0.5) < 5e-3
    z0, z1 = px.box_muller(x[0], x[1])
    z = np.concatenate([z0, z1])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


def test_partition_independence():
    a = px.synthetic_images(6, 32, 32, seed=3)
    b = px.synthetic_images(3, 32, 32, seed=3, first_image=3)
    assert np.array_equal(a[3:], b)
    ya = C.corrupt(a, "gaussian_noise", 3, seed=5)
    yb = C.corrupt(b, "gaussian_noise", 3, seed=5, first_image=3)
    assert np.array_equal(ya[3:], yb)


@pytest.mark.parametrize("name", list(C.GENERATORS))
def test_corruption_range_and_determinism(name):
    x = px.synthetic_images(3, 32, 32, seed=1)
    for s in (1, 5):
        y = C.corrupt(x, name, s, seed=2)
        assert y.dtype == np.float32 and y.shape == x.shape
        assert y.min() >= 0 and y.max() <= 1
        assert np.array_equal(y, C.corrupt(x, name, s, seed=2))
    d1 = np.abs(C.corrupt(x, name, 1, seed=2) - x / 255.0).mean()
    d5 = np.abs(C.corrupt(x, name, 5, seed=2) - x / 255.0).mean()
    assert d1 > 0 and d5 > 0
    if name not in ("glass_blur", "frost", "elastic_transform"):       # (on i.i.d. images these are not monotone in severity)
        assert d5 > d1          # severity monotone on i.i.d. images


def _probe_images(h, w, rng):
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = np.stack([(xx * 4) % 256, (yy * 3 + xx) % 256, ((xx - w // 2) ** 2 + (yy - h // 2) ** 2) // 8 % 256], -1)
    smooth = np.clip(smooth + rng.integers(-8, 9, smooth.shape), 0, 255).astype(np.uint8)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    sat = np.where(rng.random((h, w, 3)) < 0.5, 0, 255).astype(np.uint8)
    return np.stack([smooth, noise, sat])


def test_jpeg_restatement_is_byte_exact_against_pillow():
    """PINNED: oracle/jpeg.py (libjpeg restated stage by stage) == PIL save(quality) + reload, byte for byte, on smooth,
    noisy and saturated images, every quality of both severity tables, MCU-aligned and ragged sizes (even and odd)."""
    from oracle import jpeg as J
    rng = np.random.default_rng(0)
    qualities = sorted(set(C.CONSTANTS["cifar"]["jpeg_compression"] + C.CONSTANTS["imagenet"]["jpeg_compression"])) + [1, 95, 100]
    for (h, w) in ((32, 32), (224, 224), (48, 64), (33, 47), (40, 24), (18, 9), (120, 160)):
        x = _probe_images(h, w, rng)
        for q in (qualities if h * w <= 64 * 64 else qualities[::3]):
            assert np.array_equal(J.jpeg_roundtrip_u8(x, q), J.pil_roundtrip_u8(x, q)), (h, w, q)
    x = px.synthetic_images(2, 32, 32, seed=4)
    assert np.array_equal(np.rint(C.corrupt(x, "jpeg_compression", 3) * 255).astype(np.uint8), J.pil_roundtrip_u8(x, 58))


def test_pixelate_restatement_is_byte_exact_against_pillow():
    """PINNED: the Pillow BOX resampler restated in oracle/corruptions.py == Image.resize(BOX) down + up, byte for byte."""
    rng = np.random.default_rng(1)
    for (h, w) in ((32, 32), (224, 224), (120, 160), (33, 47)):
        x = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        prof = C.profile_for(h, w)
        for s in range(1, 6):
            c = C.CONSTANTS[prof]["pixelate"][s - 1]
            assert np.array_equal(np.rint(C.pixelate(x, s) * 255).astype(np.uint8), C.pixelate_pil(x, c)), (h, w, s)


def test_shot_noise_is_poisson():
    x = np.full((2, 64, 64, 3), 64, np.uint8)
    y = C.shot_noise(x, 1, profile="imagenet")          # c = 60, lambda = 15.06
    k = y * 60
    lam = 64 / 255 * 60
    assert abs(k.mean() - lam) < 0.1 and abs(k.var() - lam) < 0.5


def test_brightness_matches_colorsys():
    import colorsys
    x = px.synthetic_images(1, 8, 8, seed=4)
    x[0, 0, 0] = 0
    y = C.brightness(x, 3, profile="imagenet")[0]
    for i in range(8):
        for j in range(8):
            r, g, b = (x[0, i, j] / 255.0).tolist()
            h, s, v = colorsys.rgb_to_hsv(r, g, b)
            want = colorsys.hsv_to_rgb(h, s, min(v + 0.3, 1.0))
            assert np.allclose(y[i, j], want, atol=1e-6)


def test_defocus_matches_cv2_filter2d():
    import cv2
    x = px.synthetic_images(1, 32, 32, seed=6)
    k = C.disk_kernel(1.5, 0.1)
    want = np.stack([cv2.filter2D(x[0, ..., c].astype(np.float32) / 255, -1, k) for c in range(3)], -1)
    got = C.defocus_blur(x, 5, profile="cifar")[0]
    assert np.abs(got - np.clip(want, 0, 1)).max() < 1e-5


def test_resnet_restatement_matches_torchvision():
    import torch
    for model, hw, ncls in (("resnet18", 32, 10), ("resnet50", 64, 1000)):
        net = M.build_torchvision(model, ncls, 0)
        fd = M.fold_resnet(net)
        x = np.random.default_rng(0).standard_normal((2, hw, hw, 3)).astype(np.float32)
        got = M.forward(fd, x, T=1)[:, 0]
        want = net(torch.from_numpy(x).permute(0, 3, 1, 2)).detach().numpy()
        assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


def test_param_and_mac_counts():
    net = M.build_torchvision("resnet18", 10, 0)
    assert sum(p.numel() for p in net.parameters()) == 11_181_642
    assert M.count_macs(M.fold_resnet(net), 32, 32) == (2_408_448, 34_608_128)


def test_dropout_mask_rate_and_t1_identity():
    m = M.dropout_mask(4, 0, 3, 2, 4096, 0.2, 0)
    assert M.dropout_threshold(0.2) == 51 and M.dropout_threshold(0.5) == 128 and M.dropout_threshold(0.0) == 0
    assert abs((m == 0).mean() - 51 / 256) < 0.01
    assert np.all(m[m > 0] == np.float32(256.0) / np.float32(205.0))       # exact inverse of the realised keep probability
    assert abs(m.mean() - 1.0) < 0.02                                      # unbiased
    # the sixteen channels of a chunk read sixteen DIFFERENT bytes of the chunk's Philox call
    assert sorted(zip(M._DROP_WORD.tolist(), M._DROP_BYTE.tolist())) == [(w, b) for w in range(4) for b in range(4)]
    # neighbouring channels are uncorrelated
    k = (m > 0).astype(np.float64)
    assert abs(np.corrcoef(k[:, :-1].ravel(), k[:, 1:].ravel())[0, 1]) < 0.02
    net = M.build_torchvision("resnet18", 10, 0)
    fd = M.fold_resnet(net)
    x = np.random.default_rng(1).standard_normal((2, 32, 32, 3)).astype(np.float32)
    a = M.forward(fd, x, T=3, p=0.3, seed=1)
    assert a.shape == (2, 3, 10) and not np.allclose(a[:, 0], a[:, 1])
    b = M.forward(fd, x[1:], T=3, p=0.3, seed=1, first_image=1)
    assert np.allclose(a[1:], b, atol=1e-5)


def test_uncertainty_definitions():
    rng = np.random.default_rng(0)
    z = rng.standard_normal((50, 6, 10)).astype(np.float32) * 3
    y = rng.integers(0, 10, 50).astype(np.int32)
    u = U.uncertainty(z, y, tau=0.4)
    p = np.exp(z - z.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    pb = p.mean(1)
    assert np.allclose(u["confidence"], pb.max(-1), atol=1e-6)
    assert np.array_equal(u["pred"], pb.argmax(-1))
    H = -(pb * np.log(pb)).sum(-1)
    assert np.allclose(u["entropy"], H, atol=1e-5)
    mi = H - (-(p * np.log(p)).sum(-1)).mean(1)
    assert np.allclose(u["mutual_information"], mi, atol=1e-5) and (u["mutual_information"] >= 0).all()
    assert np.array_equal(u["failure_flag"], ((pb.argmax(-1) != y) & (pb.max(-1) >= 0.4)).astype(np.uint8))
    one = U.uncertainty(z[:, :1], y)
    assert np.allclose(one["mutual_information"], 0, atol=1e-6)


def test_bucketed_auroc_matches_sklearn_and_ece():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(1)
    n, C_ = 5000, 10
    conf = rng.uniform(0.1, 1.0, n).astype(np.float32)
    H = rng.uniform(0, 2.3, n).astype(np.float32)
    mi = rng.uniform(0, 1.0, n).astype(np.float32)
    labels = rng.integers(0, C_, n).astype(np.int32)
    pred = np.where(rng.uniform(size=n) < conf, labels, (labels + 1) % C_).astype(np.int32)
    ar = np.zeros(X.arena_words(C_), np.int64)
    X.accumulate(ar, conf, H, mi, pred, labels, 0.9, C_)
    out = X.finalize(ar, C_)
    wrong = (pred != labels).astype(int)
    s0, s1, s2 = X.normalised_scores(conf, H, mi, C_)
    for nm, s in (("auroc_msp", s0), ("auroc_entropy", s1), ("auroc_mi", s2)):
        q = X.bucket(s)
        assert abs(out[nm] - roc_auc_score(wrong, q)) < 1e-12
        assert abs(out[nm] - roc_auc_score(wrong, s)) < 2.0 / X.N_BUCKETS
    # ECE by direct definition (right-closed bins)
    ece = 0.0
    for b in range(15):
        m = (conf > np.float32(b) / 15) & (conf <= np.float32(b + 1) / 15)
        if m.any():
            ece += m.mean() * abs((pred[m] == labels[m]).mean() - conf[m].astype(np.float64).mean())
    assert abs(out["ece"] - ece) < 1e-6
    assert out["n"] == n and abs(out["accuracy"] - (pred == labels).mean()) < 1e-12
    # order independence / shard additivity (the multi-GPU reduction is a plain integer sum)
    a1 = np.zeros_like(ar); a2 = np.zeros_like(ar)
    X.accumulate(a1, conf[:2000], H[:2000], mi[:2000], pred[:2000], labels[:2000], 0.9, C_)
    X.accumulate(a2, conf[2000:], H[2000:], mi[2000:], pred[2000:], labels[2000:], 0.9, C_)
    assert np.array_equal(a1 + a2, ar)


# ------------------------------------------------------------------------------------------- f4: trust replay (PINNED)
def test_trust_replay_oracle_reproduces_the_reference_engine():
    """oracle/trust.py against trajectories of the REAL TrustEngine (tests/golden/make_golden_trust.py): the five float64
    state variables bit for bit, policy / contradiction outputs exactly."""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_trust_inputs import sequences
    from oracle import trust as OT
    with open(os.path.join(os.path.dirname(__file__), "golden", "trust_replay.json")) as fh:
        gold = json.load(fh)
    for c in gold["cases"]:
        status, score = sequences(c["seed"], c["n_seq"], c["n_ticks"])
        res = OT.replay(status, score, c["dt"])
        t = np.array(c["trajectories"], dtype=np.float64)
        assert np.array_equal(res["state"], t[:, :, :5])
        assert np.array_equal(res["policy"], t[:, :, 5]) and np.array_equal(res["contradiction"], t[:, :, 6])
        assert np.array_equal(res["contradiction_count"], t[:, :, 7])
        assert t[:, :, 6].sum() > 0                       # the contradiction branch is exercised
        for s_, i_ in ((0, 0), (1, c["n_ticks"] // 2), (c["n_seq"] - 1, c["n_ticks"] - 1)):
            st = OT.reference_state(res, s_, i_)
            assert st["reliability"] == t[s_, i_, 8] and st["trust_velocity"] == t[s_, i_, 9]


# ---------------------------------------------------------------- third-party anchors of the corruption building blocks
# (the reference ships no corruption code: these pin the restated pieces to the libraries the canonical ImageNet-C /
# CIFAR-10-C generators call -- scipy.ndimage and PIL -- at the versions in this image)
def test_gaussian_taps_match_scipy_gaussian_filter():
    import scipy.ndimage as ndi
    for sigma in (0.4, 0.7, 1.0, 1.5, 3.0):
        r, k = C.gaussian_taps(sigma)
        imp = np.zeros(2 * r + 9)
        imp[r + 4] = 1.0
        g = ndi.gaussian_filter1d(imp, sigma, mode="constant")          # truncate = 4.0, as skimage.filters.gaussian
        assert np.abs(g[4:4 + 2 * r + 1] - k).max() < 1e-15 and g[:4].max() == 0.0


def test_zoom_resampling_matches_scipy_zoom():
    import scipy.ndimage as ndi
    rng = np.random.default_rng(3)
    for size in (32, 224):
        x = (rng.integers(0, 256, (size, size, 3), dtype=np.uint8).astype(np.float32) / np.float32(255)).astype(np.float32)
        for z in (1.06, 1.11, 1.33):
            hc, top, ho, trim = C.zoom_geometry(size, z)
            ref = ndi.zoom(x[top:top + hc, top:top + hc], (z, z, 1), order=1)      # make_imagenet_c.clipped_zoom
            assert ref.shape[0] == ho
            ref = ref[trim:trim + size, trim:trim + size]
            i0, i1, fr = C._zoom_sample_axis(size, z)
            fy, fx = fr[:, None, None], fr[None, :, None]
            t = x[i0][:, i0] * (1 - fx) + x[i0][:, i1] * fx
            b = x[i1][:, i0] * (1 - fx) + x[i1][:, i1] * fx
            assert np.abs(t * (1 - fy) + b * fy - ref).max() < 5e-5


def test_pixelate_tracks_pil_box_resize():
    from PIL import Image
    rng = np.random.default_rng(4)
    for size in (32, 224):
        prof = C.profile_for(size, size)
        x = rng.integers(0, 256, (2, size, size, 3), dtype=np.uint8)
        for sev in range(1, 6):
            c = C.CONSTANTS[prof]["pixelate"][sev - 1]
            small = int(size * c)
            ref = np.stack([np.asarray(Image.fromarray(x[i]).resize((small, small), Image.BOX).resize((size, size), Image.BOX))
                            for i in range(2)])
            d = np.abs(np.rint(C.pixelate(x, sev) * 255.0).astype(np.int64) - ref.astype(np.int64))
            # PIL weights partially covered source pixels, the restatement takes whole-pixel boxes: never more than one code apart
            assert d.max() <= 1
