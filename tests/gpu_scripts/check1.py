"""First-contact GPU check (dev script, run under gpurun): calibration, NTT and MSM parity
against the oracle on small sizes, rough timings on the benchmark sizes."""
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cap_b200 import device, field  # noqa: E402
from oracle import bn254, msm as omsm, ntt as ontt  # noqa: E402

out = {}
ctx = device.Context(0)
out["calib"] = ctx.calibrate()
print("calibrate", out["calib"], flush=True)

rng = random.Random(11)
# ---- NTT parity
bad = 0
for log_n in [1, 3, 8, 10, 11, 12, 14]:
    n = 1 << log_n
    for in_len in sorted({n, min(n, n // 8 + 3)}):
        x = [rng.randrange(bn254.R) for _ in range(in_len)]
        xm = field.fr_to_mont_array(x)
        for inverse in (False, True):
            for coset in (False, True):
                got = field.fr_from_mont_array(ctx.ntt(xm, log_n, inverse, coset))
                fn = {(False, False): ontt.fft, (False, True): ontt.coset_fft, (True, False): ontt.ifft, (True, True): ontt.coset_ifft}[(inverse, coset)]
                exp = fn(x, log_n)
                ok = got == exp
                if not ok:
                    bad += 1
                    nd = sum(1 for a, b in zip(got, exp) if a != b)
                    print(f"NTT MISMATCH log_n={log_n} in_len={in_len} inv={inverse} coset={coset} ndiff={nd}", flush=True)
print("ntt parity mismatches:", bad, flush=True)
out["ntt_bad"] = bad

# batch NTT
x = [[rng.randrange(bn254.R) for _ in range(1 << 11)] for _ in range(3)]
got = ctx.ntt(np.stack([field.fr_to_mont_array(v) for v in x]), 11, False, True)
okb = all(field.fr_from_mont_array(got[i]) == ontt.coset_fft(x[i], 11) for i in range(3))
print("ntt batch ok:", okb, flush=True)
out["ntt_batch_ok"] = okb

# ---- MSM parity
tau = 0x1F2E3D4C5B6A79881726354453627180ABCDEF0123456789
mbad = 0
for npts, wb in [(1, 0), (7, 0), (300, 0), (300, 5), (1000, 11), (5000, 0)]:
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([tau])[0], size=npts, window_bits=wb)
    pts = field.g1_from_mont_array(srs.export())
    if npts <= 300:
        exp_pts = bn254.srs_powers(tau, npts) if npts > 7 else [bn254.g1_mul(bn254.G1_GEN, pow(tau, i, bn254.R)) for i in range(npts)]
        if pts != exp_pts:
            mbad += 1
            print("SRS MISMATCH", npts, flush=True)
    cases = {
        "random": [rng.randrange(bn254.R) for _ in range(npts)],
        "zeros": [0] * npts,
        "ones": [1] * npts,
        "max": [bn254.R - 1] * npts,
        "small": [rng.randrange(4) for _ in range(npts)],
        "same": [12345678901234567890] * npts,
    }
    for name, sc in cases.items():
        exp = omsm.kzg_commit_tau(sc, tau)
        for mont in (True, False):
            arr = field.fr_to_mont_array(sc) if mont else field.fr_raw_array(sc)
            got = field.g1_from_mont_array(srs.msm(arr, mont=mont))[0]
            if got != exp:
                mbad += 1
                print(f"MSM MISMATCH n={npts} wb={wb} case={name} mont={mont}", flush=True)
    if npts == 300:
        # arkworks-shaped Pippenger on the exported bases, and a batched call
        sc = cases["random"]
        assert omsm.msm_arkworks(pts, sc) == omsm.kzg_commit_tau(sc, tau)
        b = np.stack([field.fr_to_mont_array(cases["random"]), field.fr_to_mont_array(cases["small"]), field.fr_to_mont_array(cases["zeros"])])
        gotb = field.g1_from_mont_array(srs.msm(b))
        expb = [omsm.kzg_commit_tau(cases[k], tau) for k in ("random", "small", "zeros")]
        if gotb != expb:
            mbad += 1
            print("MSM BATCH MISMATCH", flush=True)
        # base offset
        got = field.g1_from_mont_array(srs.msm(field.fr_to_mont_array(sc[:100]), base_off=17))[0]
        exp = omsm.msm_naive(pts[17:117], sc[:100])
        if got != exp:
            mbad += 1
            print("MSM OFFSET MISMATCH", flush=True)
    srs.close()
print("msm parity mismatches:", mbad, flush=True)
out["msm_bad"] = mbad

# ---- rough timings (host-buffer API, includes copies)
def timeit(f, reps=5):
    f()
    ctx.sync()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        ctx.sync()
        ts.append(time.perf_counter() - t)
    return min(ts)

for log_n in (15, 18, 20):
    a = np.random.default_rng(1).integers(0, 1 << 60, size=(7, 1 << log_n, 4), dtype=np.uint64)
    out[f"ntt_{log_n}_x7_ms"] = timeit(lambda: ctx.ntt(a, log_n, False, True)) * 1e3
    print("ntt", log_n, out[f"ntt_{log_n}_x7_ms"], flush=True)
for log_n in (15, 17):
    n = (1 << log_n) + 3
    t0 = time.perf_counter()
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([tau])[0], size=n)
    out[f"srs_setup_{log_n}_s"] = time.perf_counter() - t0
    sc = np.random.default_rng(2).integers(0, 1 << 62, size=(5, n, 4), dtype=np.uint64)
    sc[..., 3] &= (1 << 60) - 1
    out[f"msm_{log_n}_x5_ms"] = timeit(lambda: srs.msm(sc, mont=False)) * 1e3
    out[f"msm_{log_n}_x1_ms"] = timeit(lambda: srs.msm(sc[0], mont=False)) * 1e3
    # spot-check correctness at full size through p(tau)
    one = [int.from_bytes(sc[0, i].tobytes(), "little") for i in range(n)]
    got = field.g1_from_mont_array(srs.msm(sc[0], mont=False))[0]
    out[f"msm_{log_n}_ok"] = got == omsm.kzg_commit_tau(one, tau)
    print("msm", log_n, out[f"msm_{log_n}_x5_ms"], out[f"msm_{log_n}_x1_ms"], out[f"msm_{log_n}_ok"], out[f"srs_setup_{log_n}_s"], flush=True)
    srs.close()
out["launches"] = ctx.launch_count
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/check1.json", "w"), indent=1)
print(json.dumps(out))
