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
    assert lib.fav_abi_version() == 2
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


def _tap_entries(ip, tab):
    """Decode the tap-list table format of csrc/tables.cu pack_taps: [(dys, dxs, ws)] per entry."""
    n_entries, max_taps = ip[0], ip[1]
    rec = 16 + 8 * max_taps
    out = []
    for i in range(n_entries):
        o = i * rec
        nt = int(tab[o:o + 4].view(np.int32)[0])
        t = tab[o + 16:o + 16 + 8 * nt].view(np.uint32).reshape(nt, 2)
        dys = (t[:, 0] & 0xFFFF).astype(np.uint16).view(np.int16).astype(int).tolist()
        dxs = (t[:, 0] >> 16).astype(np.uint16).view(np.int16).astype(int).tolist()
        out.append((dys, dxs, np.ascontiguousarray(t[:, 1]).view(np.float32)))
    return out


def test_library_constants_match_oracle_and_spec():
    """The 15 x 5 x 2 severity grid exists three times -- C++ product (tables.cu), Python display copy (spec.SEVERITY) and the
    oracle (CONSTANTS): all equal."""
    from fav import _lib, spec
    assert spec.CORRUPTIONS == OC.CORRUPTIONS and spec.CORRUPTION_ID == OC.CORRUPTION_ID
    for prof in ("cifar", "imagenet"):
        assert spec.MEAN_STD[prof] == OC.MEAN_STD[prof]
        for name in spec.IMPLEMENTED:
            assert spec.SEVERITY[prof][name] == OC.CONSTANTS[prof][name]
            for sev in range(1, 6):
                want = OC.CONSTANTS[prof][name][sev - 1]
                want = [float(v) for v in (want if isinstance(want, tuple) else (want,))]
                assert _lib.corruption_constants(prof, spec.CORRUPTION_ID[name], sev) == want, (prof, name, sev)


def test_library_host_tables_match_oracle_definitions():
    """fav_corrupt_params (host only, csrc/tables.cu) against the oracle's independent numpy / scipy / cv2 / PIL statements
    of the same tables: Poisson thresholds (scipy), disk kernel (cv2.GaussianBlur), motion taps, zoom geometry, libjpeg
    quantisation tables, glass-blur fixed-point taps, impulse thresholds."""
    from fav import _lib, spec
    from oracle import jpeg as OJ
    ID = spec.CORRUPTION_ID
    # shot noise
    for prof, sev in (("cifar", 1), ("cifar", 4), ("imagenet", 1), ("imagenet", 5)):
        fp, ip, tab = _lib.corrupt_params(ID["shot_noise"], sev, 32, 32, prof)
        c = OC.CONSTANTS[prof]["shot_noise"][sev - 1]
        k2, w2, t2 = OC.poisson_table(c)
        assert fp == [float(c)] and ip[0] == w2
        assert np.array_equal(tab[:1024].view(np.int32), k2)
        t1 = tab[1024:1024 + 1024 * w2].view(np.uint32).reshape(256, w2)
        assert np.abs(t1.astype(np.int64) - t2.astype(np.int64)).max() <= 1      # fp64 cdf rounding: at most 1 LSB of 2^-32
        jump = tab[1024 + 1024 * w2:1024 + 1024 * w2 + 131072].view(np.uint16).reshape(256, 256)
        edges = (np.arange(256, dtype=np.uint64) << np.uint64(24))
        assert np.array_equal(jump, (t1[:, None, :].astype(np.uint64) < edges[None, :, None]).sum(-1))
        # guide table of k1_shot_smem: cells b >= 1 hold the k of the draw b << 24 and the number of thresholds inside the
        # cell (count up); cell 0 holds the k at the cell's END and the number of non-zero thresholds inside (count down)
        assert ip[1] > 0 and ip[1] % 16 == 0
        guide = tab[ip[1]:ip[1] + 131072].view(np.uint16).reshape(256, 256)
        inside = np.diff(np.concatenate([jump.astype(np.int64), np.full((256, 1), w2)], 1), axis=1)
        assert np.array_equal((guide & 1023)[:, 1:], (k2[:, None] + jump)[:, 1:])
        assert np.array_equal((guide >> 10)[:, 1:], np.minimum(inside, 63)[:, 1:])
        assert np.array_equal(guide[:, 0] & 1023, k2 + jump[:, 1])
        nonzero0 = ((t1 > 0) & (t1 < (1 << 24))).sum(1)
        assert np.array_equal(guide[:, 0] >> 10, np.minimum(nonzero0, 63))
        # the search k1_shot_smem runs on it (three predicated probes, then the rare loop), emulated in numpy, returns the
        # plain inverse-CDF count for every (value, draw) -- including draws of 0, 2^32 - 1 and exact cell boundaries
        rng = np.random.default_rng(sev)
        n = 100000
        v = rng.integers(0, 256, n)
        u = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        u[:500], u[500:1000] = 0xFFFFFFFF, 0
        u[1000:2000] = rng.integers(0, 256, 1000).astype(np.uint32) << 24
        u[2000:12000] = rng.integers(0, 1 << 24, 10000).astype(np.uint32)                 # the lower-tail cell
        want = k2[v] + (t1[v] <= u[:, None]).sum(1)
        gg = guide[v, u >> 24]
        k = (gg & 1023).astype(np.int64)
        cnt = (gg >> 10).astype(np.int64)
        idx = k - k2[v]
        down = (u >> 24) == 0
        step = np.zeros(n, np.int64)
        for i in range(3):
            ok = cnt > i
            pos = np.clip(np.where(down, idx - 1 - i, idx + i), 0, w2 - 1)
            t = t1[v, pos]
            step += ok & np.where(down, t > u, t <= u)
        got = np.where(down, k - step, k + step)
        for j in np.nonzero((cnt > 3) & (step == 3))[0]:
            if down[j]:
                pp = idx[j] - 4
                while pp >= 0 and t1[v[j], pp] > u[j]:
                    pp -= 1
                    got[j] -= 1
            else:
                pp = idx[j] + 3
                while pp < w2 and t1[v[j], pp] <= u[j]:
                    pp += 1
                    got[j] += 1
        assert np.array_equal(got, want), (prof, sev, int((got != want).sum()))
    # impulse thresholds
    for prof in ("cifar", "imagenet"):
        for sev in range(1, 6):
            _, ip, _ = _lib.corrupt_params(ID["impulse_noise"], sev, 32, 32, prof)
            tp, ts = OC.impulse_thresholds(OC.CONSTANTS[prof]["impulse_noise"][sev - 1])
            assert [v & 0xFFFFFFFF for v in ip] == [tp, ts]
    # defocus: nonzero taps of the cv2 disk kernel in row-major order
    for prof in ("cifar", "imagenet"):
        for sev in range(1, 6):
            _, ip, tab = _lib.corrupt_params(ID["defocus_blur"], sev, 224, 224, prof)
            (dys, dxs, ws), = _tap_entries(ip, tab)
            dy2, dx2, w2 = OC.defocus_taps(*OC.CONSTANTS[prof]["defocus_blur"][sev - 1])
            assert dys == dy2 and dxs == dx2 and np.abs(ws - np.asarray(w2)).max() < 1e-7 and ip[2] == 0
    # motion: 91 integer angles, taps cut where the shift leaves the frame
    for (prof, sev, h, w) in (("imagenet", 5, 224, 224), ("cifar", 5, 32, 32), ("imagenet", 4, 12, 40)):
        _, ip, tab = _lib.corrupt_params(ID["motion_blur"], sev, h, w, prof)
        ent = _tap_entries(ip, tab)
        radius, sigma = OC.CONSTANTS[prof]["motion_blur"][sev - 1]
        assert len(ent) == OC.MOTION_ANGLES and ip[2] == 1
        for ai in (0, 17, 45, 60, 90):
            dy2, dx2, w2 = OC.motion_taps(radius, sigma, ai - 45)
            keep = next((j for j, (a, b) in enumerate(zip(dy2, dx2)) if abs(a) >= h or abs(b) >= w), len(w2))
            assert ent[ai][0] == dy2[:keep] and ent[ai][1] == dx2[:keep] and np.allclose(ent[ai][2], w2[:keep], atol=1e-7)
    # zoom: factors (numpy.arange quirks included) and the clipped-zoom sampling geometry
    for (prof, sev, h, w) in (("cifar", 3, 32, 32), ("imagenet", 5, 224, 224), ("imagenet", 1, 120, 160)):
        _, ip, tab = _lib.corrupt_params(ID["zoom_blur"], sev, h, w, prof)
        zs = OC.zoom_factors(OC.CONSTANTS[prof]["zoom_blur"][sev - 1])
        assert ip == [len(zs)]
        t = tab.view(np.uint32).reshape(len(zs), h + w, 2)
        for i, z in enumerate(zs):
            for off, size in ((0, h), (h, w)):
                j0, j1, gr = OC._zoom_sample_axis(size, z)
                assert np.array_equal(t[i, off:off + size, 0] & 0xFFFF, j0) and np.array_equal(t[i, off:off + size, 0] >> 16, j1)
                assert np.array_equal(np.ascontiguousarray(t[i, off:off + size, 1]).view(np.float32), gr)
    # jpeg: libjpeg quantisation tables
    for prof in ("cifar", "imagenet"):
        for sev in range(1, 6):
            _, ip, tab = _lib.corrupt_params(ID["jpeg_compression"], sev, 32, 32, prof)
            ql, qc = OJ.quant_tables(OC.CONSTANTS[prof]["jpeg_compression"][sev - 1])
            assert np.array_equal(tab.view(np.int32), np.concatenate([ql.ravel(), qc.ravel()]))
    # glass blur: exact fixed-point taps (the swapped bytes depend on them bit for bit) + fp32 taps
    for prof in ("cifar", "imagenet"):
        for sev in range(1, 6):
            sigma, delta, iters = OC.CONSTANTS[prof]["glass_blur"][sev - 1]
            _, ip, tab = _lib.corrupt_params(ID["glass_blur"], sev, 32, 32, prof)
            r, q = OC.gaussian_taps_q16(sigma)
            assert ip == [delta, iters, r]
            assert np.array_equal(tab[:4 * (2 * r + 1)].view(np.int32), q)
            assert np.allclose(tab[4 * (2 * r + 1):].view(np.float32), OC.gaussian_taps(sigma)[1].astype(np.float32), atol=1e-9)
    # snow: tap lists at -135..-45 degrees + zoom geometry
    _, ip, tab = _lib.corrupt_params(ID["snow"], 3, 32, 32, "cifar")
    loc, scale, zoom, thresh, mb_r, mb_s, blend = OC.CONSTANTS["cifar"]["snow"][2]
    ent = _tap_entries(ip, tab)
    dy2, dx2, w2 = OC.motion_taps(int(mb_r), float(mb_s), 30 - 135)
    assert ent[30][0] == dy2 and ent[30][1] == dx2 and np.allclose(ent[30][2], w2, atol=1e-7)
    z = tab[ip[7]:].view(np.uint32).reshape(64, 2)
    j0, j1, gr = OC._zoom_sample_axis(32, float(zoom))
    assert np.array_equal(z[:32, 0] & 0xFFFF, j0) and np.array_equal(np.ascontiguousarray(z[:32, 1]).view(np.float32), gr)


def test_library_pixelate_table_reproduces_pil():
    """The BOX coefficient table the pixelate kernels consume, applied in numpy exactly as k1_pixelate_down / _up do,
    equals Pillow's own resize(BOX) down + up byte for byte."""
    from fav import _lib, spec
    rng = np.random.default_rng(3)
    for (prof, h, w) in (("cifar", 32, 32), ("imagenet", 224, 224), ("imagenet", 120, 160), ("cifar", 33, 47)):
        x = rng.integers(0, 256, (1, h, w, 3), dtype=np.uint8)
        for sev in range(1, 6):
            c = OC.CONSTANTS[prof]["pixelate"][sev - 1]
            _, (sw, sh, kh, kv), tab = _lib.corrupt_params(spec.CORRUPTION_ID["pixelate"], sev, h, w, prof)
            t = tab.view(np.int32)
            hx = t[:sw * (2 + kh)].reshape(sw, 2 + kh)
            vy = t[sw * (2 + kh):sw * (2 + kh) + sh * (2 + kv)].reshape(sh, 2 + kv)
            upx = t[sw * (2 + kh) + sh * (2 + kv):][:w]
            upy = t[sw * (2 + kh) + sh * (2 + kv) + w:][:h]
            img = x[0].astype(np.int64)
            tmp = np.empty((h, sw, 3), np.int64)
            for i in range(sw):
                xmin, cnt = hx[i, 0], hx[i, 1]
                tmp[:, i] = np.clip(((img[:, xmin:xmin + cnt] * hx[i, 2:2 + cnt][None, :, None]).sum(1) + (1 << 21)) >> 22, 0, 255)
            small = np.empty((sh, sw, 3), np.int64)
            for j in range(sh):
                ymin, cnt = vy[j, 0], vy[j, 1]
                small[j] = np.clip(((tmp[ymin:ymin + cnt] * vy[j, 2:2 + cnt][:, None, None]).sum(0) + (1 << 21)) >> 22, 0, 255)
            got = small[upy][:, upx].astype(np.uint8)
            assert np.array_equal(got, OC.pixelate_pil(x, c)[0]), (prof, h, w, sev)


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


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "signal_analyzer.py")), reason="baseline/_ref not installed (run build())")
def test_bench_reference_arm_c5_times_the_real_reference_code():
    """bench.py --impl reference --config C5: the reference's own SignalAnalyzer.analyze_frame + TrustEngine.update per frame,
    imported from the unmodified copy in baseline/_ref (kind 'reference'), p50 latency in ms, lower is better."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C5", "--steps", "50"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["p99_ms"] >= line["value"]


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


def test_shot_noise_exact_division_by_fma_correction():
    """k1_shot_smem computes k / c as q = k * rc, q' = fma(fma(-q, c, k), rc, q): equal to the correctly rounded fp32 quotient for
    every k the guide table can hold and every c of both severity tables."""
    from fractions import Fraction
    f32 = np.float32

    def rn(x):
        cand = f32(float(x))
        best = None
        for d in (np.nextafter(cand, f32(-np.inf)), cand, np.nextafter(cand, f32(np.inf))):
            err = abs(Fraction(float(d)) - x)
            if best is None or err < best[0] or (err == best[0] and (int(f32(d).view(np.uint32)) & 1) == 0):
                best = (err, d)
        return f32(best[1])

    for c in sorted(set(OC.CONSTANTS["cifar"]["shot_noise"] + OC.CONSTANTS["imagenet"]["shot_noise"])):
        cf = f32(c)
        rc = f32(1.0) / cf
        for k in range(0, 1024, 1 if c in (3, 500) else 7):
            kf = f32(k)
            q = kf * rc
            e = rn(Fraction(float(-q)) * Fraction(float(cf)) + Fraction(float(kf)))
            assert rn(Fraction(float(e)) * Fraction(float(rc)) + Fraction(float(q))) == kf / cf, (c, k)


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
    """Host table of elastic_transform for long Gaussians (tables.cu elastic_fold): one weight per (destination, source) pixel.
    Against the oracle's tap-by-tap fp32 accumulation the displacement differs by far less than a thousandth of a pixel."""
    from fav import _lib
    from oracle import corruptions as K
    for (h, w, sev) in ((224, 224, 1), (32, 32, 2)):
        fp, ip, tab = _lib.corrupt_params(K.CORRUPTION_ID["elastic_transform"], sev, h, w)
        assert ip[1] == 1
        assert abs(fp[0] - np.float32(K.elastic_params(h, w, K.CONSTANTS[K.profile_for(h, w)]["elastic_transform"][sev - 1])[0])) < 1e-4
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
    fp, ip, tab = _lib.corrupt_params(K.CORRUPTION_ID["elastic_transform"], 5, 224, 224)
    assert ip[1] == 0
    r, k = K.elastic_gauss_taps(K.elastic_params(224, 224, K.CONSTANTS["imagenet"]["elastic_transform"][4])[1])
    assert ip[0] == r and np.allclose(tab.view(np.float32), k, atol=1e-9)
    # ImageNet-C scales by the literal 244 on 224-pixel frames: sigma = 0.7 * 244 at severity 1
    assert _lib.corrupt_params(K.CORRUPTION_ID["elastic_transform"], 1, 224, 224)[1][0] == int(3 * 0.7 * 244 + 0.5)


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


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _reference_live_tick():
    """The live-mode body of the reference's loop (platform/backend/main.py:153-196: analyze -> engine.update -> state keys
    -> attribution -> logging), lifted as TEXT from the unmodified copy in baseline/_ref/main.py into a plain function
    (main.py itself is never imported)."""
    import textwrap
    lines = open(os.path.join(REF_DIR, "main.py")).read().splitlines()
    live = "\n".join(lines[152:188])                 # :153-188, body of `else:  # Live mode`
    tail = "\n".join(lines[189:196])                 # :190-196, `if state:` attribution + log
    assert "analyzer.analyze_frame(frame)" in live and "logger.log(state" in tail, "reference main.py changed"
    src = ("def tick(video_src, analyzer, engine, attributor, logger, source_mode, dt, last_processed_frame_id, last_analysis,\n"
           "         _frame_to_base64_jpeg):\n    state = None\n" + textwrap.indent(textwrap.dedent(live), "    ") + "\n" +
           textwrap.indent(textwrap.dedent(tail), "    ") + "\n    return state, last_processed_frame_id, last_analysis\n")
    ns = {}
    exec(compile(src, "baseline/_ref/main.py:153-196", "exec"), ns)
    return ns["tick"]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "main.py")), reason="baseline/_ref not installed (run build())")
def test_gate_dict_flows_through_the_reference_loop():
    """f3: the dict UncertaintyGate.analyze_frame returns (SignalAnalyzer's keys + metrics['uncertainty']) goes through the
    reference's own loop body unchanged: engine.update accepts it, the state payload stays JSON-serialisable (ws.send_json,
    main.py:200), SessionLogger.log and FailureAttributor.update work, and the extra keys reach state['signal_metrics']."""
    sys.path.insert(0, REF_DIR)
    try:
        from trust_engine import TrustEngine
        from session_logger import SessionLogger
        from failure_attributor import FailureAttributor
    finally:
        sys.path.remove(REF_DIR)
    from fav.gate import SignalFinisher, assemble_result
    from oracle import frame_stats as OFS
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_frames import frame_sequence

    class StubGate:
        """UncertaintyGate without the GPU: same SignalFinisher + assemble_result, frame statistics from the CPU oracle,
        classifier outputs faked."""
        def __init__(self):
            self.fin, self.prev = SignalFinisher(), None

        def analyze_frame(self, frame):
            st, self.prev = OFS.frame_stats(frame, self.prev)
            conf = 0.5 + 0.4 * float(frame[0, 0, 0]) / 255.0
            return assemble_result(self.fin.finish(st, frame.shape[0] * frame.shape[1]), (conf, 2.5, 0.3, 17), "uncertainty", 0.9, 1000)

    class Src:
        def __init__(self, frames):
            self.frames, self.i = frames, 0

        def get_frame(self):
            self.i += 1
            return self.frames[(self.i - 1) % len(self.frames)], self.i

    tick = _reference_live_tick()
    engine, logger, attributor, gate = TrustEngine(), SessionLogger(), FailureAttributor(), StubGate()
    src = Src(frame_sequence(0, 96, 128))
    last_id, last_analysis = 0, None
    for _ in range(12):
        state, last_id, last_analysis = tick(src, gate, engine, attributor, logger, "video", 1 / 30, last_id, last_analysis,
                                             lambda f: "jpeg")
        assert state is not None
        payload = json.loads(json.dumps(state))                              # what ws.send_json would put on the wire
        u = payload["signal_metrics"]["uncertainty"]
        assert set(u) == {"confidence", "entropy", "mutual_information", "pred", "normalized_entropy", "high_confidence"}
        assert payload["anomaly_score"] == round(2.5 / np.log(1000), 6) and payload["vision_status"].startswith("VISION_")
        assert 0.0 <= payload["reliability"] <= 1.0 and "failure_events" in payload
    rows = logger.get_csv().strip().splitlines()
    assert len(rows) == 13 and rows[0].split(",") == SessionLogger.HEADER


def test_sweep_wire_format_records():
    """f3: per-cell records (CSV in session_logger.py style, JSON like the 'sequence_result' reply of main.py:354-357) carry the
    metrics plus device time, evals/s and roofline fraction; NaN AUROC (single-class cell) becomes null / empty."""
    from fav.sweep import CorruptionSweep
    res = {("fog", 1): {"n": 10, "accuracy": 0.5, "ece": 0.123456789, "mean_confidence": 0.3, "mean_entropy": 1.0,
                        "mean_mutual_information": 0.0, "failure_rate": 0.1, "auroc_msp": float("nan"), "auroc_entropy": 0.6,
                        "auroc_mi": 0.5}}
    perf = {("fog", 1): {"gpu_ms": 1.5, "evals_per_gpu_s": 6666.6, "tflops": 100.0, "roofline_frac": 0.07}}
    doc = json.loads(CorruptionSweep.to_json(res, perf, {"n_gpus": 2}))
    assert doc["type"] == "sweep_result" and doc["n_gpus"] == 2
    (c,) = doc["cells"]
    assert c["auroc_msp"] is None and c["ece"] == 0.123457 and c["roofline_frac"] == 0.07 and c["evals_per_gpu_s"] == 6666.6
    hdr, row = CorruptionSweep.to_csv(res, perf).strip().splitlines()
    assert hdr.split(",") == CorruptionSweep.COLUMNS + CorruptionSweep.PERF_COLUMNS
    assert row.split(",")[:3] == ["fog", "1", "10"] and row.split(",")[9] == ""


def test_work_items_interleave_the_grid():
    """Any 20 consecutive steps (the driver's bench window) visit all 15 corruptions and all 5 severities, any 15 at least 14
    corruptions; per-rank round-robin shares sample the whole grid; every (cell, block) item appears exactly once."""
    from fav.sweep import CorruptionSweep, SweepConfig

    class _NoGpu(CorruptionSweep):
        def __init__(self, cfg):
            self.cfg, self.cells = cfg, cfg.cells()

    sw = _NoGpu(SweepConfig(block=100))
    items = sw.work_items(250)
    assert sorted(items) == [(ci, b) for ci in range(75) for b in range(3)]
    for start in (0, 7, 60, 75, 100):
        win = items[start:start + 15]
        assert len({sw.cells[ci].name for ci, _ in win}) >= 14
        assert len({sw.cells[ci].severity for ci, _ in win}) == 5
        assert len({sw.cells[ci].name for ci, _ in items[start:start + 20]}) == 15
    for world in (2, 4, 8):                       # a rank's round-robin share of the first 75 items still mixes the grid
        share = [sw.cells[items[i][0]].name for i in range(0, 75, world)]
        assert len(set(share)) >= 0.75 * min(15, len(share))


def test_graft_entry_build_is_consistent():
    """The driver's build check: __graft_entry__.build() compiles the library (incremental), installs baseline/_ref and asserts
    the ABI version the header declares."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "fav_b200.h")).read()
    from fav import _lib
    assert int(re.search(r"#define FAV_ABI_VERSION (\d+)", hdr).group(1)) == _lib.load().fav_abi_version()
