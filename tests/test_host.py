"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host logic (spec tables,
weight packing, partitioning, finalisation, multi-rank reduction over gloo) agrees with the oracle."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import corruptions as OC
from oracle import metrics as OX
from oracle import model as OM
from oracle import philox as px
from oracle import uncertainty as OU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from fav import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "fav_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fav_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fav_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert lib.fav_abi_version() == 1
    assert lib.fav_hist_words(10, 15, 4096) == OX.arena_words(10)
    assert lib.fav_hist_words(1000, 15, 4096) == OX.arena_words(1000)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_fallback():
    import fav
    from fav import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.fav_init(0, ctypes.byref(h)) != 0 and lib.fav_last_error()
    with pytest.raises(RuntimeError):
        fav.VisionClassifier()
    with pytest.raises(RuntimeError):
        fav.UncertaintyGate()


def test_spec_matches_oracle_tables():
    from fav import spec
    assert spec.CORRUPTIONS == OC.CORRUPTIONS and spec.CORRUPTION_ID == OC.CORRUPTION_ID
    for prof in ("cifar", "imagenet"):
        for name in spec.IMPLEMENTED:
            assert spec.SEVERITY[prof][name] == OC.CONSTANTS[prof][name]
        assert spec.MEAN_STD[prof] == OC.MEAN_STD[prof]
    for c in (500, 75, 60, 3):
        k1, w1, t1 = spec.poisson_table(c)
        k2, w2, t2 = OC.poisson_table(c)
        assert w1 == w2 and np.array_equal(k1, k2)
        assert np.abs(t1.astype(np.int64) - t2.astype(np.int64)).max() <= 1      # fp64 cdf rounding at most 1 LSB of 2^-32
    for r, a in ((0.3, 0.4), (1.5, 0.1), (3, 0.1), (8, 0.5), (10, 0.5)):
        assert np.abs(spec.disk_kernel(r, a) - OC.disk_kernel(r, a)).max() < 1e-7
    for ang in (-45, -7, 0, 33, 45):
        a = spec.motion_taps(15, 8, ang, 224, 224)
        b = OC.motion_taps(15, 8, ang)
        assert a[0] == b[0] and a[1] == b[1] and np.allclose(a[2], b[2], atol=1e-7)
    lo, hi = spec._pixelate_axis(224, 0.3)
    lo2, hi2 = OC.pixelate_geometry(224, 0.3)
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    i0, i1, fr = spec._zoom_axis(32, 1.21)
    j0, j1, gr = OC._zoom_sample_axis(32, 1.21)
    assert np.array_equal(i0, j0) and np.array_equal(i1, j1) and np.array_equal(fr, gr)
    assert spec.zoom_factors((1.33, 0.03)) == OC.zoom_factors((1.33, 0.03))


def test_config_objects():
    import fav
    from fav.sweep import SweepConfig
    c = fav.CorruptionConfig("fog", 4)
    assert c.id == 10 and c.to_dict() == {"corruption": "fog", "severity": 4}
    assert fav.CorruptionConfig(None).id == 0 and fav.CorruptionConfig("clean", 3).severity == 0
    for bad in (("fog", 0), ("fog", 6), ("nope", 1)):
        with pytest.raises(ValueError):
            fav.CorruptionConfig(*bad)
    cfg = SweepConfig(include_clean=True)
    assert len(cfg.cells()) == 1 + 5 * len(fav.IMPLEMENTED) == 76          # the full 15 x 5 grid + clean
    json.dumps(cfg.to_dict())             # trivially serialisable, like the reference's JSON actions


def test_weight_blob_layout_and_folding():
    from fav import weights
    net = weights.build_model("resnet18", 10, 0, logit_gain=8.0)
    blob = weights.pack_resnet(net, "resnet18")
    assert blob[:8] == b"FAVW1\0\0\0"
    n_convs, ncls, mid, n_blocks = np.frombuffer(blob, np.int32, 4, 8)
    assert (n_convs, ncls, mid, n_blocks) == (21, 10, 18, 8)
    bt = np.frombuffer(blob, np.int32, 16, 24).reshape(8, 2)
    assert bt[:, 0].tolist() == [2] * 8 and bt[:, 1].tolist() == [0, 0, 1, 0, 1, 0, 1, 0]
    off = (24 + 64 + 15) // 16 * 16
    folded = OM.fold_resnet(OM.build_torchvision("resnet18", 10, 0, logit_gain=8.0))
    rec = np.frombuffer(blob, np.int32, 8, off)
    assert rec[:6].tolist() == [64, 3, 7, 7, 2, 3]
    w = np.frombuffer(blob, np.uint16, 64 * 147, off + 32).astype(np.uint32) << 16
    want = folded["convs"][0]["w"].permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).float().numpy().ravel()
    assert np.array_equal(w.view(np.float32), want)
    b50 = weights.pack_resnet(weights.build_model("resnet50", 1000, 0), "resnet50")
    assert np.frombuffer(b50, np.int32, 4, 8).tolist() == [54, 1000, 50, 16]


def test_partition_and_finalize():
    from fav.sweep import finalize, partition
    for n, w in ((800, 8), (7, 3), (0, 2), (5, 8)):
        parts = [partition(n, r, w) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    rng = np.random.default_rng(0)
    for C_ in (10, 1000):
        n = 4000
        z = (rng.standard_normal((n, 3, C_)) * 2).astype(np.float32)
        y = rng.integers(0, C_, n).astype(np.int32)
        u = OU.uncertainty(z, y, 0.4)
        ar = np.zeros(OX.arena_words(C_), np.int64)
        OX.accumulate(ar, u["confidence"], u["entropy"], u["mutual_information"], u["pred"], y, 0.4, C_)
        a, b = finalize(ar, C_, 15, 4096), OX.finalize(ar, C_)
        assert set(a) == set(b)
        for k in a:
            assert a[k] == b[k] or (np.isnan(a[k]) and np.isnan(b[k])), k


def _rank_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from fav.sweep import allreduce_arena, partition
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, C_, T, block, cells = 96, 10, 3, 16, [("gaussian_noise", 2), ("contrast", 5), ("impulse_noise", 1)]
    x, y = px.synthetic_images(n, 32, 32, 1), px.synthetic_labels(n, 10, 1)
    rng = np.random.default_rng(7)
    W = rng.standard_normal((3072, C_)).astype(np.float32) * 0.05
    items = [(ci, b) for b in range(n // block) for ci in range(len(cells))]
    arena = np.zeros((len(cells), OX.arena_words(C_)), np.int64)
    for i in partition(len(items), rank, world):
        ci, b = items[i]
        lo = b * block
        xc = OC.corrupt(x[lo:lo + block], cells[ci][0], cells[ci][1], seed=1, first_image=lo)
        # a stand-in linear classifier keeps the CPU test fast; the sharding / reduction logic is what is under test
        z = np.stack([(xc.reshape(block, -1) + 0.01 * t) @ W for t in range(T)], 1).astype(np.float32) * 20
        u = OU.uncertainty(z, y[lo:lo + block], 0.3)
        OX.accumulate(arena[ci], u["confidence"], u["entropy"], u["mutual_information"], u["pred"], y[lo:lo + block], 0.3, C_)
    t = torch.from_numpy(arena)
    allreduce_arena(t)
    np.save(os.path.join(tmp, f"arena_w{world}_r{rank}.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_reduction_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    tmp = str(tmp_path)
    for world, port in ((1, 29611), (2, 29612)):
        mp.start_processes(_rank_worker, args=(world, port, tmp), nprocs=world, join=True, start_method="spawn")
    one = np.load(os.path.join(tmp, "arena_w1_r0.npy"))
    r0 = np.load(os.path.join(tmp, "arena_w2_r0.npy"))
    r1 = np.load(os.path.join(tmp, "arena_w2_r1.npy"))
    assert np.array_equal(r0, r1) and np.array_equal(r0, one)
    assert one[:, 0].tolist() == [96, 96, 96]


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_trust_engine_golden_transcript_is_recorded():
    """The consumer (TrustEngine.update, trust_engine.py:139) is unchanged; the golden transcript of the
    reference's own smoke script (test_trust.py) is kept as the regression vector."""
    with open(os.path.join(ROOT, "tests", "golden", "trust_engine.json")) as fh:
        tr = json.load(fh)["transcript"]
    assert [round(r[1], 6) for r in tr] == [1.0, 0.5149, 0.0, 0.0, 0.501721]
    assert [r[2] for r in tr] == ["VISION_ALLOWED", "VISION_DEGRADED", "VISION_BLOCKED", "VISION_BLOCKED", "VISION_DEGRADED"]


def test_div255_fma_correction_is_exact():
    """corrupt.cu::div255 claims b*r + FMA residual correction == correctly rounded b/255 for every byte."""
    from fractions import Fraction
    f32 = np.float32

    def fma(a, b, c):
        x = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
        cand = f32(float(x))
        best = None
        for d in (np.nextafter(cand, f32(-np.inf)), cand, np.nextafter(cand, f32(np.inf))):
            err = abs(Fraction(float(d)) - x)
            if best is None or err < best[0] or (err == best[0] and (int(f32(d).view(np.uint32)) & 1) == 0):
                best = (err, d)
        return f32(best[1])

    r = f32(0.003921568859368563)
    assert r == f32(1.0) / f32(255.0)
    for b in range(256):
        q = f32(b) * r
        assert fma(fma(-q, f32(255.0), f32(b)), r, q) == f32(b) / f32(255.0)


def test_cli_and_trust_helpers_without_gpu():
    """`python -m fav.sweep --help` parses without touching the GPU; the trust-replay host helpers map names to codes."""
    import subprocess
    out = subprocess.run([sys.executable, "-m", "fav.sweep", "--help"], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert out.returncode == 0 and "--passes" in out.stdout and "--corruptions" in out.stdout
    from fav import trust
    codes = trust.status_codes([["VISION_OK", "VISION_BLANK"], ["VISION_CORRUPTED", "VISION_FROZEN"]])
    assert codes.dtype == np.int8 and codes.tolist() == [[0, 2], [3, 1]]
    assert trust.POLICY[3] == "VISION_BLOCKED" and trust.DECAY_RATES["VISION_BLANK"] == 0.60


def test_dropout_four_byte_compare_is_exact_for_every_threshold():
    """The device compares four Philox bytes at a time (csrc/common.cuh dropout_keep4: AND, ADD, one LOP3, then prmt.b32 sign
    replication).  Restated in numpy: the sign bit of every result byte equals (byte >= thr8) for all 256 thresholds, and the
    PRMT selectors 0x9988 / 0xBBAA turn the flags of bytes (0,1) / (2,3) into 0xFFFF-per-channel masks."""
    rng = np.random.default_rng(0)
    r = rng.integers(0, 2 ** 32, 50000, dtype=np.uint64).astype(np.uint32)
    r[:256] = np.arange(256, dtype=np.uint32) * np.uint32(0x01010101)          # every byte value in every lane
    for thr in range(256):
        add4 = np.uint32((((0x80 - thr) if thr <= 128 else (0x100 - thr)) * 0x01010101) & 0xFFFFFFFF)
        hi4 = np.uint32(0xFFFFFFFF if thr > 128 else 0)
        t = ((r & np.uint32(0x7F7F7F7F)) + add4).astype(np.uint32)
        keep4 = (t & r) | (~hi4 & (t | r))
        flags = [((keep4 >> np.uint32(8 * b + 7)) & 1).astype(bool) for b in range(4)]
        for b in range(4):
            assert np.array_equal(flags[b], ((r >> np.uint32(8 * b)) & 0xFF) >= thr), (thr, b)
        lo = np.where(flags[0], 0xFFFF, 0).astype(np.uint32) | (np.where(flags[1], 0xFFFF, 0).astype(np.uint32) << np.uint32(16))
        # prmt.b32 with selector nibble 8|k = byte k's sign replicated: 0x9988 -> [s0, s0, s1, s1]
        sign = [np.where(f, 0xFF, 0).astype(np.uint32) for f in flags]
        prmt_lo = sign[0] | (sign[0] << np.uint32(8)) | (sign[1] << np.uint32(16)) | (sign[1] << np.uint32(24))
        assert np.array_equal(prmt_lo, lo)


def test_elastic_folded_matrices_reproduce_the_tap_order_smoothing():
    """Host table of elastic_transform for long Gaussians (spec.elastic_fold): one weight per (destination, source) pixel.
    Against the oracle's tap-by-tap fp32 accumulation the displacement differs by far less than a thousandth of a pixel."""
    import fav
    from fav import spec
    from oracle import corruptions as K
    for (h, w, sev) in ((224, 224, 1), (32, 32, 2)):
        fp, ip, tab = spec.kernel_params(fav.CorruptionConfig("elastic_transform", sev), h, w)
        assert ip[1] == 1
        t = tab.view(np.float32)
        mwt, mh = t[:w * w].reshape(w, w), t[w * w:].reshape(h, h)
        assert np.allclose(mwt.sum(0), 1.0, atol=1e-5) and np.allclose(mh.sum(1), 1.0, atol=1e-5)      # a smoothing kernel
        alpha, sigma, _ = K.elastic_params(h, w, K.CONSTANTS[K.profile_for(h, w)]["elastic_transform"][sev - 1])
        r, k = K.elastic_gauss_taps(sigma)
        assert r == ip[0]
        u = (np.random.default_rng(0).random((h, w)).astype(np.float32) * 2 - 1)
        p1 = np.zeros_like(u)
        for tt in range(2 * r + 1):
            p1 += k[tt] * u[:, K._reflect_sym(np.arange(w) + tt - r, w)]
        d = np.zeros_like(u)
        for tt in range(2 * r + 1):
            d += k[tt] * p1[K._reflect_sym(np.arange(h) + tt - r, h)]
        df = mh.astype(np.float64) @ (u.astype(np.float64) @ mwt.astype(np.float64))
        assert np.abs(d - df).max() * alpha < 1e-4
    # short kernels keep the tap list
    assert spec.kernel_params(fav.CorruptionConfig("elastic_transform", 5), 224, 224)[1][1] == 0


def test_glass_swap_wavefront_schedule_equals_the_sequential_chain():
    """k1_glass runs the scan-order swap chain of glass_blur as a row wavefront: lane r walks scan row r, `lag` columns behind
    lane r-1, lag = max(2 delta, ceil(sw / 32)), bands of rows one after the other.  Simulated here step by step (all swaps of
    a step applied 'simultaneously': none may share a pixel) against the plain sequential loop of oracle/corruptions.py."""
    rng = np.random.default_rng(7)
    for (h, w, delta, iters, band) in ((12, 14, 1, 2, 5), (16, 13, 2, 3, 100), (20, 40, 4, 1, 7), (9, 70, 1, 2, 3)):
        sh, sw = h - 2 * delta, w - 2 * delta
        rows_total = iters * sh
        dy = rng.integers(-delta, delta, size=rows_total * sw)
        dx = rng.integers(-delta, delta, size=rows_total * sw)
        img0 = rng.permutation(h * w).reshape(h, w)
        seq = img0.copy()
        j = 0
        for _ in range(iters):
            for hh in range(h - delta, delta, -1):
                for ww in range(w - delta, delta, -1):
                    h2, w2 = hh + dy[j], ww + dx[j]
                    j += 1
                    seq[hh, ww], seq[h2, w2] = seq[h2, w2], seq[hh, ww]
        wav = img0.copy()
        lag = max(2 * delta, (sw + 31) // 32)
        for r0 in range(0, rows_total, band):
            nb = min(band, rows_total - r0)
            for s in range((nb - 1) * lag + sw):
                touched, swaps = set(), []
                for r in range(nb):
                    k = s - r * lag
                    if 0 <= k < sw:
                        R = r0 + r
                        hh, ww = h - delta - R % sh, w - delta - k
                        jj = R * sw + k
                        a, b = (hh, ww), (hh + dy[jj], ww + dx[jj])
                        assert a not in touched and b not in touched, "two swaps of one step share a pixel"
                        touched.update((a, b))
                        swaps.append((a, b))
                for a, b in swaps:
                    wav[a], wav[b] = wav[b], wav[a]
        assert np.array_equal(seq, wav), (h, w, delta, iters, band)
