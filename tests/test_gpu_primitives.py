"""GPU parity tests for the standalone primitives, through the C ABI (capgpu_ntt, capgpu_msm_g1):
bit-exact against the oracle at sizes it finishes in seconds, golden fixtures, and
size-independent properties at the benchmark sizes."""
import os
import random

import numpy as np
import pytest

from cap_b200 import _lib, device, field
from oracle import bn254 as B
from oracle import msm as omsm
from oracle import ntt as ontt

from conftest import TAU

pytestmark = pytest.mark.gpu

ORACLE_NTT = {(False, False): ontt.fft, (False, True): ontt.coset_fft, (True, False): ontt.ifft, (True, True): ontt.coset_ifft}


def _h(xs):
    return [int(x, 16) for x in xs]


def _pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


@pytest.mark.parametrize("log_n", [1, 2, 5, 9, 10, 11, 12, 13])
def test_ntt_matches_oracle(ctx, log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    for in_len in sorted({n, min(n, n // 8 + 3), 1}):
        x = [rng.randrange(B.R) for _ in range(in_len)]
        xm = field.fr_to_mont_array(x)
        for (inverse, coset), fn in ORACLE_NTT.items():
            got = field.fr_from_mont_array(ctx.ntt(xm, log_n, inverse, coset))
            assert got == fn(x, log_n), (log_n, in_len, inverse, coset)


def test_ntt_golden(ctx, golden):
    for v in golden["ntt"]:
        xm = field.fr_to_mont_array(_h(v["input"]))
        for (inverse, coset), key in {(False, False): "fft", (True, False): "ifft", (False, True): "coset_fft", (True, True): "coset_ifft"}.items():
            assert field.fr_from_mont_array(ctx.ntt(xm, v["log_n"], inverse, coset)) == _h(v[key])


def test_ntt_edge_values_and_batch(ctx):
    log_n = 11
    n = 1 << log_n
    zeros = field.fr_to_mont_array([0] * n)
    assert field.fr_from_mont_array(ctx.ntt(zeros, log_n)) == [0] * n
    # constant polynomial c -> all evaluations c; delta at X^1 on the coset -> 5 * omega^i
    c = B.R - 1
    assert field.fr_from_mont_array(ctx.ntt(field.fr_to_mont_array([c]), log_n)) == [c] * n
    w = B.fr_root_of_unity(log_n)
    got = field.fr_from_mont_array(ctx.ntt(field.fr_to_mont_array([0, 1]), log_n, coset=True))
    assert got[:4] == [5 * pow(w, i, B.R) % B.R for i in range(4)]
    rng = random.Random(3)
    xs = [[rng.randrange(B.R) for _ in range(n)] for _ in range(3)]
    got = ctx.ntt(np.stack([field.fr_to_mont_array(v) for v in xs]), log_n, inverse=True, coset=True)
    for i in range(3):
        assert field.fr_from_mont_array(got[i]) == ontt.coset_ifft(xs[i], log_n)


@pytest.mark.parametrize("log_n", [15, 18, 20])
def test_ntt_round_trip_and_linearity_at_bench_sizes(ctx, log_n):
    """Size-independent properties at the prover's sizes (2^15 = n, 2^18 = 8n, 2^20 = largest)."""
    n = 1 << log_n
    g = np.random.default_rng(log_n)
    a = g.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 60) - 1  # Montgomery limbs of some value < r
    for coset in (False, True):
        f = ctx.ntt(a, log_n, False, coset)
        back = ctx.ntt(f, log_n, True, coset)
        assert np.array_equal(back, a)
    # spot-check a few outputs against direct evaluation of the polynomial (first 64 coefficients non-zero)
    short = a[:64]
    coeffs = field.fr_from_mont_array(short)
    f = field.fr_from_mont_array(ctx.ntt(short, log_n, False, True)[[0, 1, n // 2 + 3, n - 1]])
    w = B.fr_root_of_unity(log_n)
    for got, i in zip(f, [0, 1, n // 2 + 3, n - 1]):
        assert got == ontt.poly_eval(coeffs, 5 * pow(w, i, B.R) % B.R)


def _rho(log_n):
    """Generator of the 3 * 2^log_n-point quotient domain: 5^((r-1) / (3 * 2^log_n)), rho^3 = the 2^log_n-th root of unity."""
    rho = pow(5, (B.R - 1) // (3 << log_n), B.R)
    assert pow(rho, 3, B.R) == B.fr_root_of_unity(log_n)
    return rho


@pytest.mark.parametrize("log_n", [3, 7, 10, 11, 13])
def test_ntt3_matches_oracle(ctx, log_n):
    """capgpu_ntt3_dev against the oracle's coset transforms with the shifts 5 rho^k (one- and two-pass sizes), and its inverse
    against the coefficients of a random polynomial of degree < 3N evaluated through the oracle."""
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    rho = _rho(log_n)
    shifts = [5 * pow(rho, k, B.R) % B.R for k in range(3)]
    for in_len in sorted({n, n // 2 + 3, 1}):
        x = [rng.randrange(B.R) for _ in range(in_len)]
        got = ctx.ntt3(field.fr_to_mont_array(x)[None], log_n)[0]
        for k in range(3):
            assert field.fr_from_mont_array(got[k]) == ontt.coset_fft(x, log_n, shifts[k]), (log_n, in_len, k)
    assert shifts == ontt.domain3_shifts(log_n)
    # f = T_0 + X^N T_1 + X^2N T_2 on the coset s_k H: sum_a (s_k^N)^a T_a(s_k w^i)
    coeffs = [rng.randrange(B.R) for _ in range(3 * n)]
    vals = []
    for k in range(3):
        c = pow(shifts[k], n, B.R)
        parts = [ontt.coset_fft(coeffs[a * n:(a + 1) * n], log_n, shifts[k]) for a in range(3)]
        vals.append([(parts[0][i] + c * parts[1][i] + c * c % B.R * parts[2][i]) % B.R for i in range(n)])
    vm = np.stack([field.fr_to_mont_array(v) for v in vals])[None]
    assert field.fr_from_mont_array(ctx.ntt3(vm, log_n, inverse=True)[0]) == coeffs
    if log_n <= 10:  # the oracle's own restatement of the transform pair (O(n) modular powers per coefficient: small sizes)
        assert vals == ontt.fft3(coeffs, log_n) and ontt.ifft3(vals, log_n) == coeffs
    # batch of two, rows independent
    two = np.concatenate([vm, vm[:, ::-1]])
    back = ctx.ntt3(two, log_n, inverse=True)
    assert field.fr_from_mont_array(back[0]) == coeffs and not np.array_equal(back[0], back[1])


def test_ntt3_at_the_prover_size(ctx):
    """Size-independent checks at 2n = 2^16 (the transfer circuit's quotient domain): direct evaluation of a few points, and
    inverse(forward) on a polynomial of the quotient's degree 5n + 7 assembled from three forward transforms."""
    log_n = 16
    n = 1 << log_n
    g = np.random.default_rng(16)
    a = g.integers(0, 1 << 62, size=(3, n, 4), dtype=np.uint64)
    a[:, :, 3] &= (1 << 60) - 1
    a[2, n // 2 + 8:] = 0  # degree 5 * (n / 2) + 7
    rho, w = _rho(log_n), B.fr_root_of_unity(log_n)
    f = ctx.ntt3(a, log_n)  # (3 chunks, 3 cosets, n, 4)
    short = a[0, :48].copy()
    fs = ctx.ntt3(short[None], log_n)[0]
    cs = field.fr_from_mont_array(short)
    for k, i in [(0, 0), (1, 1), (2, n // 2 + 3), (2, n - 1)]:
        x = 5 * pow(rho, k, B.R) * pow(w, i, B.R) % B.R
        assert field.fr_from_mont_array(fs[k, i:i + 1])[0] == ontt.poly_eval(cs, x)
    vals = np.empty((3, n, 4), dtype=np.uint64)
    for k in range(3):
        c = pow(5 * pow(rho, k, B.R) % B.R, n, B.R)
        t = [field.fr_from_mont_array(f[ch, k]) for ch in range(3)]
        vals[k] = field.fr_to_mont_array([(t[0][i] + c * t[1][i] + c * c % B.R * t[2][i]) % B.R for i in range(n)])
    back = ctx.ntt3(vals[None], log_n, inverse=True)[0]
    assert np.array_equal(back, a.reshape(3 * n, 4))


def test_ntt_argument_errors(ctx):
    a = np.zeros((4, 4), dtype=np.uint64)
    with pytest.raises(_lib.CapGpuError):
        ctx.ntt(a, 1)  # input longer than the domain
    with pytest.raises(_lib.CapGpuError):
        ctx.ntt(a, 21)  # unsupported size
    with pytest.raises(_lib.CapGpuError):
        ctx.ntt3(a[None], 1)  # input longer than 2^log_n


@pytest.fixture(scope="module")
def small_srs(ctx):
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=300)
    yield srs
    srs.close()


def test_srs_setup_matches_oracle(ctx, small_srs, golden):
    pts = field.g1_from_mont_array(small_srs.export(40))
    assert pts == [_pt(p) for p in golden["msm"]["srs"]]
    assert pts[:8] == B.srs_powers(TAU, 8)


def test_msm_golden_and_upload_path(ctx, golden):
    g = golden["msm"]
    srs_pts = [_pt(p) for p in g["srs"]]
    srs = device.Srs(ctx, points_xy=field.g1_to_mont_array(srs_pts))
    for c in g["cases"]:
        sc = _h(c["scalars"])
        assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc)))[0] == _pt(c["result"]), c["name"]
        assert field.g1_from_mont_array(srs.msm(field.fr_raw_array(sc), mont=False))[0] == _pt(c["result"]), c["name"]
    srs.close()


@pytest.mark.parametrize("window_bits", [0, 2, 5, 10, 11, 13, 16])
def test_msm_edge_cases(ctx, window_bits):
    n = 300
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n, window_bits=window_bits)
    rng = random.Random(window_bits)
    cases = {
        "random": [rng.randrange(B.R) for _ in range(n)],
        "zeros": [0] * n,
        "ones": [1] * n,
        "max": [B.R - 1] * n,
        "small": [rng.randrange(4) for _ in range(n)],
        "same": [0x1234567890ABCDEF] * n,
        "half": [(B.R - 1) // 2 + rng.randrange(3) for _ in range(n)],
    }
    for name, sc in cases.items():
        exp = omsm.kzg_commit_tau(sc, TAU)
        assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc)))[0] == exp, name
    # ragged: shorter vectors, base offsets, a batch
    sc = cases["random"]
    pts = field.g1_from_mont_array(srs.export())
    assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc[:1])))[0] == B.g1_mul(pts[0], sc[0])
    assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc[:50]), base_off=250))[0] == omsm.msm_naive(pts[250:300], sc[:50])
    b = np.stack([field.fr_to_mont_array(cases[k]) for k in ("random", "small", "zeros", "max", "ones")])
    assert field.g1_from_mont_array(srs.msm(b)) == [omsm.kzg_commit_tau(cases[k], TAU) for k in ("random", "small", "zeros", "max", "ones")]
    with pytest.raises(_lib.CapGpuError) as e:
        srs.msm(field.fr_to_mont_array(sc[:51]), base_off=250)
    assert e.value.code == -4  # CAPGPU_ERR_SRS_TOO_SMALL
    srs.close()


@pytest.mark.parametrize("log_n", [13, 15])
def test_msm_skewed_scalars_use_the_heavy_bucket_path(ctx, log_n):
    """Witness-like scalar vectors: mostly 0 / 1 / small values and a repeated constant, so a few
    buckets hold thousands of entries (one CTA per heavy bucket) while the rest stay light or empty.
    At 2^15 the lone-MSM schedule takes the flat (equal chunks of sorted entries) accumulation,
    whose chunks then start and end inside, at and across bucket boundaries."""
    n = 1 << log_n
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
    rng = random.Random(77)
    const = rng.randrange(B.R)
    vecs = []
    for frac in (0.3, 0.9, 1.0):
        sc = []
        for _ in range(n):
            u = rng.random()
            if u < frac * 0.5:
                sc.append(rng.randrange(2))
            elif u < frac * 0.8:
                sc.append(rng.randrange(256))
            elif u < frac:
                sc.append(const)
            else:
                sc.append(rng.randrange(B.R))
        vecs.append(sc)
    exp = [omsm.kzg_commit_tau(v, TAU) for v in vecs]
    assert field.g1_from_mont_array(srs.msm(np.stack([field.fr_to_mont_array(v) for v in vecs]))) == exp
    for v, e in zip(vecs, exp):  # each vector alone: the lone-MSM schedule
        assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(v)))[0] == e
    srs.close()


def test_msm_repeated_bases(ctx):
    """Buckets that receive P, P and -P exercise the doubling / cancellation branches."""
    P = B.g1_mul(B.G1_GEN, 99)
    pts = [P, P, B.g1_neg(P), None] * 16
    srs = device.Srs(ctx, points_xy=field.g1_to_mont_array(pts), window_bits=4)
    sc = [7] * 64
    assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc)))[0] == B.g1_mul(P, 7 * 16)
    sc = [3, 3, 6, 1] * 16
    assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc)))[0] is None
    srs.close()


@pytest.mark.parametrize("log_n", [15, 17])
def test_msm_at_bench_sizes(ctx, log_n):
    """Full-size check through an independent route: sum s_i tau^i G = p(tau) G, and linearity."""
    n = (1 << log_n) + 3
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
    g = np.random.default_rng(log_n)
    a = g.integers(0, 1 << 62, size=(2, n, 4), dtype=np.uint64)
    a[..., 3] &= (1 << 60) - 1
    res = field.g1_from_mont_array(srs.msm(a, mont=False))
    for i in range(2):
        sc = field.fr_from_raw_array(a[i])
        assert res[i] == omsm.kzg_commit_tau(sc, TAU)
    s0, s1 = field.fr_from_raw_array(a[0]), field.fr_from_raw_array(a[1])
    ssum = field.fr_raw_array([(x + y) % B.R for x, y in zip(s0, s1)])
    assert field.g1_from_mont_array(srs.msm(ssum, mont=False))[0] == B.g1_add(res[0], res[1])
    # lone-MSM schedule (flat accumulation + lane-pair reduction) on degenerate vectors: nothing to add,
    # one giant bucket, a single non-zero scalar at either end
    zeros = np.zeros((n, 4), dtype=np.uint64)
    assert field.g1_from_mont_array(srs.msm(zeros, mont=False))[0] is None
    ones = zeros.copy()
    ones[:, 0] = 1
    assert field.g1_from_mont_array(srs.msm(ones, mont=False))[0] == omsm.kzg_commit_tau([1] * n, TAU)
    for pos in (0, n - 1):
        one = zeros.copy()
        one[pos] = a[0][pos]
        assert field.g1_from_mont_array(srs.msm(one, mont=False))[0] == B.g1_mul(B.g1_mul(B.G1_GEN, pow(TAU, pos, B.R)), s0[pos])
    srs.close()


@pytest.mark.parametrize("log_n,batch", [(14, 12), (15, 9)])
def test_msm_large_batches_use_the_strip_reduction(ctx, log_n, batch):
    """Batched launches with >= 2^17 buckets in flight take the single-lane strip / per-sum form of the bucket
    reduction (msm_red_strips, msm_red_sums_lane, 32-quad plane CTAs): every vector of the batch against
    p(tau) G, with an all-zero vector, a vector of ones and a single-entry vector among them."""
    n = (1 << log_n) + 3
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
    g = np.random.default_rng(100 + log_n)
    a = g.integers(0, 1 << 62, size=(batch, n, 4), dtype=np.uint64)
    a[..., 3] &= (1 << 60) - 1
    a[1] = 0
    a[2] = 0
    a[2, :, 0] = 1
    a[3] = 0
    a[3, n - 1] = a[0, n - 1]
    res = field.g1_from_mont_array(srs.msm(a, mont=False))
    for i in range(batch):
        assert res[i] == omsm.kzg_commit_tau(field.fr_from_raw_array(a[i]), TAU), i
    srs.close()


def test_srs_upload_from_compressed_points(ctx):
    """ark-serialize compressed G1 (the on-disk form of the reference's SRS / key files,
    src/parameters.rs:557-592): decompression on the device reproduces the points, including
    infinity and both y signs; a non-curve x is rejected."""
    from oracle.transcript import g1_compressed
    rng = random.Random(8)
    pts = [B.g1_mul(B.G1_GEN, rng.randrange(B.R)) for _ in range(20)] + [None, B.G1_GEN, B.g1_neg(B.G1_GEN)]
    blob = b"".join(g1_compressed(p) for p in pts)
    srs = device.Srs(ctx, compressed=blob, window_bits=5)
    assert field.g1_from_mont_array(srs.export()) == pts
    sc = [rng.randrange(B.R) for _ in pts]
    assert field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc)))[0] == omsm.msm_naive(pts, sc)
    srs.close()
    # x = 0 gives y^2 = 3, a quadratic non-residue mod q: not a curve point
    assert pow(3, (B.Q - 1) // 2, B.Q) == B.Q - 1
    with pytest.raises(_lib.CapGpuError):
        device.Srs(ctx, compressed=bytes(32))


@pytest.mark.parametrize("log_n,parts", [(9, 2), (12, 8), (15, 8), (6, 4)])
def test_msm_bucket_range_slices_sum_to_the_msm(ctx, log_n, parts):
    """capgpu_msm_g1_dev_part: the bucket-range slices of one MSM (what each GPU of a split MSM
    computes) add up to the whole MSM -- uniform scalars, and skewed ones that put every digit into
    the first slice (small scalars) or the last (magnitudes near 2^(c-1))."""
    import torch
    from ctypes import c_void_p
    n = (1 << log_n) + 3
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
    rng = np.random.default_rng(log_n)
    uniform = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    uniform[:, 3] &= (1 << 60) - 1
    small = np.zeros((n, 4), dtype=np.uint64)
    small[:, 0] = rng.integers(0, 4, size=n, dtype=np.uint64)
    for sc in (uniform, small):
        want = srs.msm(sc, mont=False)
        d = torch.from_numpy(sc.view(np.int64)).cuda()
        outs = torch.zeros((parts, 8), dtype=torch.int64, device="cuda")
        for p in range(parts):
            _lib.check(ctx.lib.capgpu_msm_g1_dev_part(ctx.h, srs.h, 0, c_void_p(d.data_ptr()), n, 0, p, parts, c_void_p(outs[p].data_ptr())), ctx.h)
        total = torch.zeros(8, dtype=torch.int64, device="cuda")
        _lib.check(ctx.lib.capgpu_g1_sum_dev(ctx.h, c_void_p(outs.data_ptr()), parts, c_void_p(total.data_ptr())), ctx.h)
        ctx.sync()
        assert np.array_equal(total.cpu().numpy().view(np.uint64), want)
        # the same slices left in XYZZ form (128 B each) and folded with one conversion to affine
        xy = torch.zeros((parts, 16), dtype=torch.int64, device="cuda")
        rcs = [ctx.lib.capgpu_msm_g1_dev_part_xyzz(ctx.h, srs.h, 0, c_void_p(d.data_ptr()), n, 0, p, parts, c_void_p(xy[p].data_ptr())) for p in range(parts)]
        if log_n >= 12:
            assert rcs == [0] * parts
            total2 = torch.zeros(8, dtype=torch.int64, device="cuda")
            _lib.check(ctx.lib.capgpu_g1_sum_xyzz_dev(ctx.h, c_void_p(xy.data_ptr()), parts, c_void_p(total2.data_ptr())), ctx.h)
            ctx.sync()
            assert np.array_equal(total2.cpu().numpy().view(np.uint64), want)
            # fused exchange on one GPU: every slice stores its sum into "its" slot of a (here local) buffer and releases
            # the slot's flag; the fold waits on all flags (what each GPU of a split does with its peers' slices)
            import ctypes
            buf = torch.zeros(16 * parts + 16 * ((parts + 15) // 16), dtype=torch.int64, device="cuda")
            flags_off = 16 * parts * 8
            for epoch in (1, 2):
                for p in range(parts):
                    slots = (ctypes.c_void_p * 1)(buf.data_ptr() + 128 * p)
                    flags = (ctypes.c_void_p * 1)(buf.data_ptr() + flags_off + 8 * p)
                    _lib.check(ctx.lib.capgpu_msm_g1_dev_part_peer(ctx.h, srs.h, 0, c_void_p(d.data_ptr()), n, 0, p, parts, slots, flags, 1, epoch), ctx.h)
                total3 = torch.zeros(8, dtype=torch.int64, device="cuda")
                _lib.check(ctx.lib.capgpu_g1_sum_xyzz_wait_dev(ctx.h, c_void_p(buf.data_ptr()), c_void_p(buf.data_ptr() + flags_off), 8, parts, epoch,
                                                               c_void_p(total3.data_ptr())), ctx.h)
                ctx.sync()
                assert np.array_equal(total3.cpu().numpy().view(np.uint64), want)
            assert ctx.lib.capgpu_msm_g1_dev_part_peer(ctx.h, srs.h, 0, c_void_p(d.data_ptr()), n, 0, 0, parts, None, None, 1, 1) == -2
        else:
            assert rcs == [-2] * parts  # fewer than 512 buckets per slice: affine slices only
    assert ctx.lib.capgpu_msm_g1_dev_part(ctx.h, srs.h, 0, c_void_p(d.data_ptr()), n, 0, 3, 3, c_void_p(total.data_ptr())) == -2
    srs.close()


def test_split_msm_across_gpus():
    """The split MSM on every GPU of the box (torchrun, NCCL all-gather of the slice results
    enqueued on the context stream): equals p(tau) G.  Skipped with fewer than 2 GPUs."""
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ngpu}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "gpu_scripts", "split_msm.py"), "--points", str(1 << 14), "--reps", "3"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert '"correct": true' in res.stdout
