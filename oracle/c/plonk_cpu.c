/* CPU restatement of the reference's proving hot path, in C -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Restates, algorithm for algorithm, what jf-cap 0.0.4 executes on the host cores below
 * `PlonkKzgSnark::prove` (/root/reference/src/proof/transfer.rs:181, mint.rs:113, freeze.rs:151):
 *   - ark-ff 0.3.0 Fp256 Montgomery arithmetic (4 x 64-bit limbs, binary-Euclid inverse);
 *   - ark-ec 0.3.0 Jacobian G1 (add_assign_mixed, double_in_place, add_assign) and
 *     VariableBaseMSM::multi_scalar_mul (unsigned c-bit windows, c = ln_without_floats(N)+2,
 *     2^c-1 buckets, zero/one fast paths, one task per window);
 *   - ark-poly 0.3.0 Radix2EvaluationDomain fft / ifft / coset variants;
 *   - jf-relation 0.1.2 compute_prod_permutation_polynomial (serial, one division per row);
 *   - jf-plonk 0.1.2 Prover rounds 1-5 (25 coset FFTs + point-wise quotient with one division
 *     per point, split + mask, evaluations, linearisation, batched openings by long division)
 *     with the SolidityTranscript (Keccak-256).
 * The source of those crates is NOT vendored in /root/reference (Cargo.toml:14-47), so this file
 * follows their published algorithms [UPSTREAM-RECALL, SURVEY.md App. A]; PARITY UNPINNED
 * against upstream bytes.  It is pinned against the Python big-int oracle (tests/test_c_oracle.py)
 * and is what bench.py times as the same-host CPU baseline ("port"): rayon's parallel
 * iterators are restated as pthread parallel-for loops over the same index spaces.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------ */
/* field arithmetic                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { u64 v[4]; } fe;
typedef struct { u64 p[4]; u64 inv; u64 r2[4]; u64 one[4]; } field_t;

static const field_t FR = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull},
    {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}};
static const field_t FQ = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full},
    {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}};

static inline int big_geq(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline int big_is_zero(const u64* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int big_is_one(const u64* a) { return a[0] == 1 && (a[1] | a[2] | a[3]) == 0; }
static inline void big_sub(u64* a, const u64* b) {
  u128 br = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - br; a[i] = (u64)t; br = (t >> 64) & 1; }
}
static inline u64 big_add(u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; a[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static inline void big_div2(u64* a, u64 top) {
  a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63); a[2] = (a[2] >> 1) | (a[3] << 63); a[3] = (a[3] >> 1) | (top << 63);
}

static inline fe f_add(const field_t* F, fe a, fe b) { big_add(a.v, b.v); if (big_geq(a.v, F->p)) big_sub(a.v, F->p); return a; }
static inline fe f_sub(const field_t* F, fe a, fe b) {
  if (!big_geq(a.v, b.v)) big_add(a.v, F->p);
  big_sub(a.v, b.v);
  return a;
}
static inline fe f_neg(const field_t* F, fe a) { fe z; memset(&z, 0, sizeof z); return big_is_zero(a.v) ? a : f_sub(F, z, a); }
static inline fe f_dbl(const field_t* F, fe a) { return f_add(F, a, a); }

static inline fe f_mul(const field_t* F, fe a, fe b) {
  u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  for (int i = 0; i < 4; i++) {
    u128 c = (u128)a.v[0] * b.v[i] + t0; t0 = (u64)c; c >>= 64;
    c += (u128)a.v[1] * b.v[i] + t1; t1 = (u64)c; c >>= 64;
    c += (u128)a.v[2] * b.v[i] + t2; t2 = (u64)c; c >>= 64;
    c += (u128)a.v[3] * b.v[i] + t3; t3 = (u64)c; c >>= 64;
    c += t4; t4 = (u64)c; u64 t5 = (u64)(c >> 64);
    u64 m = t0 * F->inv;
    c = (u128)m * F->p[0] + t0; c >>= 64;
    c += (u128)m * F->p[1] + t1; t0 = (u64)c; c >>= 64;
    c += (u128)m * F->p[2] + t2; t1 = (u64)c; c >>= 64;
    c += (u128)m * F->p[3] + t3; t2 = (u64)c; c >>= 64;
    c += t4; t3 = (u64)c; t4 = t5 + (u64)(c >> 64);
  }
  fe r = {{t0, t1, t2, t3}};
  if (t4 || big_geq(r.v, F->p)) big_sub(r.v, F->p);
  return r;
}
static inline fe f_sqr(const field_t* F, fe a) { return f_mul(F, a, a); }
static inline fe f_one(const field_t* F) { fe r; memcpy(r.v, F->one, 32); return r; }
static inline fe f_zero(void) { fe r; memset(&r, 0, sizeof r); return r; }
static inline int f_is_zero(fe a) { return big_is_zero(a.v); }
static inline int f_eq(fe a, fe b) { return memcmp(a.v, b.v, 32) == 0; }
static inline fe f_from_mont(const field_t* F, fe a) { fe o = {{1, 0, 0, 0}}; return f_mul(F, a, o); }
static inline fe f_to_mont(const field_t* F, fe a) { fe r2; memcpy(r2.v, F->r2, 32); return f_mul(F, a, r2); }
static inline fe f_from_u64(const field_t* F, u64 x) { fe a = {{x, 0, 0, 0}}; return f_to_mont(F, a); }

static fe f_pow_u64(const field_t* F, fe a, u64 e) {
  fe r = f_one(F);
  for (int i = 63; i >= 0; i--) { r = f_sqr(F, r); if ((e >> i) & 1) r = f_mul(F, r, a); }
  return r;
}

/* ark-ff Fp256::inverse: binary extended Euclid (Guajardo-Kumar-Paar-Pelzl Alg. 16); b starts at
 * R^2 so the result is already in Montgomery form.  inverse(0) is undefined (callers check). */
static fe f_inv(const field_t* F, fe a) {
  u64 u[4], v[4];
  fe b, c = f_zero();
  memcpy(u, a.v, 32); memcpy(v, F->p, 32); memcpy(b.v, F->r2, 32);
  while (!big_is_one(u) && !big_is_one(v)) {
    while (!(u[0] & 1)) {
      big_div2(u, 0);
      if (!(b.v[0] & 1)) big_div2(b.v, 0); else { u64 cy = big_add(b.v, F->p); big_div2(b.v, cy); }
    }
    while (!(v[0] & 1)) {
      big_div2(v, 0);
      if (!(c.v[0] & 1)) big_div2(c.v, 0); else { u64 cy = big_add(c.v, F->p); big_div2(c.v, cy); }
    }
    if (!big_geq(u, v)) { big_sub(v, u); c = f_sub(F, c, b); } else { big_sub(u, v); b = f_sub(F, b, c); }
  }
  return big_is_one(u) ? b : c;
}

/* ------------------------------------------------------------------------------------------ */
/* G1 Jacobian (ark-ec short_weierstrass_jacobian, a = 0)                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { fe x, y; } g1a;      /* all-zero == infinity (ABI convention) */
typedef struct { fe x, y, z; } g1j;   /* z == 0 is infinity */

static inline int g1a_is_inf(const g1a* p) { return f_is_zero(p->x) && f_is_zero(p->y); }
static inline g1j g1j_inf(void) { g1j r; r.x = f_one(&FQ); r.y = f_one(&FQ); r.z = f_zero(); return r; }

static void g1j_double(g1j* p) {
  const field_t* F = &FQ;
  if (f_is_zero(p->z)) return;
  fe a = f_sqr(F, p->x), b = f_sqr(F, p->y), c = f_sqr(F, b);
  fe d = f_sub(F, f_sub(F, f_sqr(F, f_add(F, p->x, b)), a), c);
  d = f_dbl(F, d);
  fe e = f_add(F, f_dbl(F, a), a), f = f_sqr(F, e);
  fe z3 = f_dbl(F, f_mul(F, p->z, p->y));
  fe x3 = f_sub(F, f, f_dbl(F, d));
  fe c8 = f_dbl(F, f_dbl(F, f_dbl(F, c)));
  fe y3 = f_sub(F, f_mul(F, f_sub(F, d, x3), e), c8);
  p->x = x3; p->y = y3; p->z = z3;
}

static void g1j_add_mixed(g1j* p, const g1a* q) {
  const field_t* F = &FQ;
  if (g1a_is_inf(q)) return;
  if (f_is_zero(p->z)) { p->x = q->x; p->y = q->y; p->z = f_one(F); return; }
  fe z1z1 = f_sqr(F, p->z);
  fe u2 = f_mul(F, q->x, z1z1);
  fe s2 = f_mul(F, f_mul(F, q->y, p->z), z1z1);
  if (f_eq(p->x, u2) && f_eq(p->y, s2)) { g1j_double(p); return; }
  fe h = f_sub(F, u2, p->x), hh = f_sqr(F, h);
  fe i = f_dbl(F, f_dbl(F, hh));
  fe j = f_mul(F, h, i);
  fe r = f_dbl(F, f_sub(F, s2, p->y));
  fe v = f_mul(F, p->x, i);
  fe x3 = f_sub(F, f_sub(F, f_sub(F, f_sqr(F, r), j), v), v);
  fe y3 = f_sub(F, f_mul(F, r, f_sub(F, v, x3)), f_dbl(F, f_mul(F, p->y, j)));
  fe z3 = f_sub(F, f_sub(F, f_sqr(F, f_add(F, p->z, h)), z1z1), hh);
  p->x = x3; p->y = y3; p->z = z3;
}

static void g1j_add(g1j* p, const g1j* q) {
  const field_t* F = &FQ;
  if (f_is_zero(q->z)) return;
  if (f_is_zero(p->z)) { *p = *q; return; }
  fe z1z1 = f_sqr(F, p->z), z2z2 = f_sqr(F, q->z);
  fe u1 = f_mul(F, p->x, z2z2), u2 = f_mul(F, q->x, z1z1);
  fe s1 = f_mul(F, f_mul(F, p->y, q->z), z2z2), s2 = f_mul(F, f_mul(F, q->y, p->z), z1z1);
  if (f_eq(u1, u2) && f_eq(s1, s2)) { g1j_double(p); return; }
  fe h = f_sub(F, u2, u1);
  fe i = f_sqr(F, f_dbl(F, h));
  fe j = f_mul(F, h, i);
  fe r = f_dbl(F, f_sub(F, s2, s1));
  fe v = f_mul(F, u1, i);
  fe x3 = f_sub(F, f_sub(F, f_sqr(F, r), j), f_dbl(F, v));
  fe y3 = f_sub(F, f_mul(F, r, f_sub(F, v, x3)), f_dbl(F, f_mul(F, s1, j)));
  fe z3 = f_mul(F, f_sub(F, f_sub(F, f_sqr(F, f_add(F, p->z, q->z)), z1z1), z2z2), h);
  p->x = x3; p->y = y3; p->z = z3;
}

static g1a g1j_to_affine(const g1j* p) {
  const field_t* F = &FQ;
  g1a r;
  if (f_is_zero(p->z)) { memset(&r, 0, sizeof r); return r; }
  fe zi = f_inv(F, p->z), zi2 = f_sqr(F, zi);
  r.x = f_mul(F, p->x, zi2);
  r.y = f_mul(F, p->y, f_mul(F, zi2, zi));
  return r;
}

/* ------------------------------------------------------------------------------------------ */
/* parallel-for (stands in for rayon's par_iter over the same index space)                     */
/* ------------------------------------------------------------------------------------------ */
typedef void (*body_fn)(size_t lo, size_t hi, void* arg);
typedef struct { body_fn fn; size_t lo, hi; void* arg; } task_t;
static void* task_run(void* p) { task_t* t = (task_t*)p; t->fn(t->lo, t->hi, t->arg); return NULL; }

static void parallel_for(size_t n, int nthreads, body_fn fn, void* arg) {
  if (nthreads <= 1 || n <= 1) { fn(0, n, arg); return; }
  size_t nt = (size_t)nthreads < n ? (size_t)nthreads : n;
  pthread_t th[256];
  task_t tk[256];
  if (nt > 256) nt = 256;
  for (size_t t = 0; t < nt; t++) {
    tk[t].fn = fn; tk[t].arg = arg; tk[t].lo = n * t / nt; tk[t].hi = n * (t + 1) / nt;
    if (t + 1 < nt) pthread_create(&th[t], NULL, task_run, &tk[t]);
  }
  task_run(&tk[nt - 1]);
  for (size_t t = 0; t + 1 < nt; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* MSM: ark-ec 0.3.0 VariableBaseMSM::multi_scalar_mul                                         */
/* ------------------------------------------------------------------------------------------ */
static int log2_ceil(size_t x) { int l = 0; while (((size_t)1 << l) < x) l++; return l; }
static int ark_window_bits(size_t n) { return n < 32 ? 3 : log2_ceil(n) * 69 / 100 + 2; }

typedef struct { const g1a* bases; const fe* scalars; size_t n; int c; g1j* sums; } msm_arg;

static void msm_window(size_t lo, size_t hi, void* argp) {
  msm_arg* a = (msm_arg*)argp;
  const int c = a->c;
  const size_t nb = ((size_t)1 << c) - 1;
  g1j* buckets = (g1j*)malloc(nb * sizeof(g1j));
  for (size_t w = lo; w < hi; w++) {
    const int w_start = (int)w * c;
    g1j res = g1j_inf();
    for (size_t i = 0; i < nb; i++) buckets[i] = g1j_inf();
    for (size_t i = 0; i < a->n; i++) {
      const u64* s = a->scalars[i].v;
      if (big_is_zero(s)) continue;
      if (big_is_one(s)) { if (w_start == 0) g1j_add_mixed(&res, &a->bases[i]); continue; }
      /* (scalar >> w_start) % 2^c */
      int limb = w_start >> 6, sh = w_start & 63;
      u64 d = s[limb] >> sh;
      if (sh + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
      d &= ((u64)1 << c) - 1;
      if (d) g1j_add_mixed(&buckets[d - 1], &a->bases[i]);
    }
    g1j running = g1j_inf();
    for (size_t i = nb; i-- > 0;) { g1j_add(&running, &buckets[i]); g1j_add(&res, &running); }
    a->sums[w] = res;
  }
  free(buckets);
}

/* scalars canonical (BigInteger256); result affine */
static g1a msm_arkworks(const g1a* bases, const fe* scalars, size_t n, int nthreads) {
  const int c = ark_window_bits(n);
  const int nwin = (254 + c - 1) / c;
  g1j sums[128];
  msm_arg a = {bases, scalars, n, c, sums};
  parallel_for((size_t)nwin, nthreads, msm_window, &a);
  g1j total = g1j_inf();
  for (int w = nwin - 1; w >= 1; w--) {
    g1j_add(&total, &sums[w]);
    for (int k = 0; k < c; k++) g1j_double(&total);
  }
  g1j_add(&total, &sums[0]);
  return g1j_to_affine(&total);
}

/* KZG10::commit: Montgomery -> canonical (into_repr), then the MSM */
static g1a kzg_commit(const g1a* srs, const fe* coeffs, size_t n, int nthreads) {
  fe* sc = (fe*)malloc((n ? n : 1) * sizeof(fe));
  for (size_t i = 0; i < n; i++) sc[i] = f_from_mont(&FR, coeffs[i]);
  g1a r = msm_arkworks(srs, sc, n, nthreads);
  free(sc);
  return r;
}

/* ------------------------------------------------------------------------------------------ */
/* FFT: ark-poly 0.3.0 Radix2EvaluationDomain                                                  */
/* ------------------------------------------------------------------------------------------ */
static const u64 ROOT28[4] = {0x636e735580d13d9cull, 0xa22bf3742445ffd6ull, 0x56452ac01eb203d8ull, 0x1860ef942963f9e7ull};
static const u64 GEN5[4] = {0x1b0d0ef99fffffe6ull, 0xeaba68a3a32a913full, 0x47d8eb76d8dd0689ull, 0x15d0085520f5bbc3ull};

static fe omega_of(unsigned log_n) {
  fe w; memcpy(w.v, ROOT28, 32);
  for (unsigned i = 0; i < 28 - log_n; i++) w = f_sqr(&FR, w);
  return w;
}

typedef struct { fe* a; size_t n, m; const fe* tw; size_t tw_stride; } fft_arg;
static void fft_stage(size_t lo, size_t hi, void* argp) {
  fft_arg* f = (fft_arg*)argp;
  const size_t m = f->m;
  for (size_t t = lo; t < hi; t++) {
    size_t blk = t / m, k = t % m;
    size_t i0 = blk * 2 * m + k, i1 = i0 + m;
    fe x = f_mul(&FR, f->a[i1], f->tw[k * f->tw_stride]);
    fe u = f->a[i0];
    f->a[i0] = f_add(&FR, u, x);
    f->a[i1] = f_sub(&FR, u, x);
  }
}

/* in place, natural order in and out: bit-reverse then DIT butterflies (arkworks' ifft shape;
 * the forward transform is computed with the same butterfly network, result identical) */
static void fft_core(fe* a, unsigned log_n, fe omega, int nthreads) {
  const size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t r = 0;
    for (unsigned b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < r) { fe t = a[i]; a[i] = a[r]; a[r] = t; }
  }
  fe* tw = (fe*)malloc((n / 2 ? n / 2 : 1) * sizeof(fe));
  tw[0] = f_one(&FR);
  for (size_t k = 1; k < n / 2; k++) tw[k] = f_mul(&FR, tw[k - 1], omega);
  for (size_t m = 1; m < n; m <<= 1) {
    fft_arg f = {a, n, m, tw, n / (2 * m)};
    parallel_for(n / 2, (n >= 4096) ? nthreads : 1, fft_stage, &f);
  }
  free(tw);
}

static void fft_inplace(fe* a, unsigned log_n, int inverse, int coset, int nthreads) {
  const size_t n = (size_t)1 << log_n;
  fe g; memcpy(g.v, GEN5, 32);
  if (!inverse) {
    if (coset) { fe s = f_one(&FR); for (size_t i = 0; i < n; i++) { a[i] = f_mul(&FR, a[i], s); s = f_mul(&FR, s, g); } }
    fft_core(a, log_n, omega_of(log_n), nthreads);
  } else {
    fft_core(a, log_n, f_inv(&FR, omega_of(log_n)), nthreads);
    fe ninv = f_inv(&FR, f_from_u64(&FR, n));
    fe gi = f_inv(&FR, g), s = ninv;
    for (size_t i = 0; i < n; i++) { a[i] = f_mul(&FR, a[i], s); if (coset) s = f_mul(&FR, s, gi); }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Keccak-256 + SolidityTranscript (twin of oracle/transcript.py)                              */
/* ------------------------------------------------------------------------------------------ */
static u64 rol64(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(u64 st[25]) {
  static const u64 RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                             0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                             0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                             0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                             0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int r = 0; r < 24; r++) {
    u64 c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) st[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(st[x + 5 * y], ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[r];
  }
}
static void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
  u64 st[25]; memset(st, 0, sizeof st);
  size_t off = 0;
  while (len - off >= 136) { for (int i = 0; i < 17; i++) { u64 l; memcpy(&l, data + off + 8 * i, 8); st[i] ^= l; } keccak_f(st); off += 136; }
  uint8_t last[136]; memset(last, 0, sizeof last);
  memcpy(last, data + off, len - off);
  last[len - off] ^= 0x01; last[135] ^= 0x80;
  for (int i = 0; i < 17; i++) { u64 l; memcpy(&l, last + 8 * i, 8); st[i] ^= l; }
  keccak_f(st);
  memcpy(out, st, 32);
}

typedef struct { uint8_t* buf; size_t len, cap; uint8_t state[64]; } transcript_t;
static void tr_init(transcript_t* t) { t->cap = 4096; t->buf = (uint8_t*)malloc(t->cap); t->len = 0; memset(t->state, 0, 64); }
static void tr_append(transcript_t* t, const void* p, size_t n) {
  if (t->len + n + 1 > t->cap) { while (t->len + n + 1 > t->cap) t->cap *= 2; t->buf = (uint8_t*)realloc(t->buf, t->cap); }
  memcpy(t->buf + t->len, p, n); t->len += n;
}
static void tr_append_fr(transcript_t* t, fe x) { fe c = f_from_mont(&FR, x); tr_append(t, c.v, 32); }
static void tr_append_g1(transcript_t* t, const g1a* p) {
  uint8_t b[32];
  if (g1a_is_inf(p)) { memset(b, 0, 32); b[31] |= 0x40; tr_append(t, b, 32); return; }
  fe x = f_from_mont(&FQ, p->x), y = f_from_mont(&FQ, p->y), ny = f_from_mont(&FQ, f_neg(&FQ, p->y));
  memcpy(b, x.v, 32);
  if (!f_eq(y, ny) && big_geq(y.v, ny.v)) b[31] |= 0x80;
  tr_append(t, b, 32);
}
static fe tr_challenge(transcript_t* t) {
  uint8_t* in = (uint8_t*)malloc(64 + t->len + 1);
  memcpy(in, t->state, 64); memcpy(in + 64, t->buf, t->len);
  uint8_t h0[32], h1[32];
  in[64 + t->len] = 0; keccak256(in, 64 + t->len + 1, h0);
  in[64 + t->len] = 1; keccak256(in, 64 + t->len + 1, h1);
  free(in);
  memcpy(t->state, h0, 32); memcpy(t->state + 32, h1, 32);
  /* from_le_bytes_mod_order(state[..48]) */
  fe acc = f_zero(), c256 = f_from_u64(&FR, 256);
  for (int i = 47; i >= 0; i--) acc = f_add(&FR, f_mul(&FR, acc, c256), f_from_u64(&FR, t->state[i]));
  return acc;
}

/* ------------------------------------------------------------------------------------------ */
/* prover                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  u64 wires_poly_comms[5][8], prod_perm_poly_comm[8], split_quot_poly_comms[5][8], opening_proof[8], shifted_opening_proof[8];
  u64 wires_evals[5][4], wire_sigma_evals[4][4], perm_next_eval[4];
} proof_t;

static fe poly_eval(const fe* p, size_t len, fe x) {
  fe acc = f_zero();
  for (size_t i = len; i-- > 0;) acc = f_add(&FR, f_mul(&FR, acc, x), p[i]);
  return acc;
}

typedef struct { fe** polys; const size_t* lens; unsigned log_m; int nthreads_inner; } cfft_arg;
static void coset_fft_many(size_t lo, size_t hi, void* argp) {
  cfft_arg* a = (cfft_arg*)argp;
  for (size_t i = lo; i < hi; i++) fft_inplace(a->polys[i], a->log_m, 0, 1, a->nthreads_inner);
}

typedef struct {
  size_t n, m; const fe* k; fe **sel, **sig, **w; fe *z, *pi, *out; fe alpha, alpha2, beta, gamma; fe zh_inv[8]; fe wm;
} quot_arg;
static void quot_body(size_t lo, size_t hi, void* argp) {
  quot_arg* q = (quot_arg*)argp;
  const field_t* F = &FR;
  fe g; memcpy(g.v, GEN5, 32);
  fe x = f_mul(F, g, f_pow_u64(F, q->wm, lo));
  fe nn = f_from_u64(F, q->n), one = f_one(F);
  for (size_t i = lo; i < hi; i++) {
    fe w[5];
    for (int j = 0; j < 5; j++) w[j] = q->w[j][i];
    fe acc = f_add(F, q->sel[11][i], q->pi[i]);
    for (int j = 0; j < 4; j++) acc = f_add(F, acc, f_mul(F, q->sel[j][i], w[j]));
    fe w01 = f_mul(F, w[0], w[1]), w23 = f_mul(F, w[2], w[3]);
    acc = f_add(F, acc, f_mul(F, q->sel[4][i], w01));
    acc = f_add(F, acc, f_mul(F, q->sel[5][i], w23));
    for (int j = 0; j < 4; j++) { fe w2 = f_sqr(F, w[j]); fe w5 = f_mul(F, f_sqr(F, w2), w[j]); acc = f_add(F, acc, f_mul(F, q->sel[6 + j][i], w5)); }
    acc = f_sub(F, acc, f_mul(F, q->sel[10][i], w[4]));
    acc = f_add(F, acc, f_mul(F, q->sel[12][i], f_mul(F, f_mul(F, w01, w23), w[4])));
    fe r1 = q->z[i], r2 = q->z[(i + 8) % q->m];
    fe bx = f_mul(F, q->beta, x);
    for (int j = 0; j < 5; j++) {
      fe wg = f_add(F, w[j], q->gamma);
      r1 = f_mul(F, r1, f_add(F, wg, f_mul(F, q->k[j], bx)));
      r2 = f_mul(F, r2, f_add(F, wg, f_mul(F, q->beta, q->sig[j][i])));
    }
    acc = f_add(F, acc, f_mul(F, q->alpha, f_sub(F, r1, r2)));
    /* (z(x) - 1) * alpha^2 / (n * (x - 1)): one field division per point, as upstream */
    fe t2 = f_mul(F, f_mul(F, q->alpha2, f_sub(F, q->z[i], one)), f_inv(F, f_mul(F, nn, f_sub(F, x, one))));
    q->out[i] = f_add(F, f_mul(F, acc, q->zh_inv[i & 7]), t2);
    x = f_mul(F, x, q->wm);
  }
}

typedef struct { const g1a* srs; fe** polys; const size_t* lens; g1a* out; int nthreads; } commit_arg;

typedef struct { fe** polys; const size_t* lens; const fe* xs; fe* out; } eval_arg;
static void eval_body(size_t lo, size_t hi, void* argp) {
  eval_arg* e = (eval_arg*)argp;
  for (size_t i = lo; i < hi; i++) e->out[i] = poly_eval(e->polys[i], e->lens[i], e->xs[i]);
}

typedef struct { fe* dst; const fe* src; fe s; } axpy_arg;
static void axpy_body(size_t lo, size_t hi, void* argp) {
  axpy_arg* a = (axpy_arg*)argp;
  for (size_t i = lo; i < hi; i++) a->dst[i] = f_add(&FR, a->dst[i], f_mul(&FR, a->src[i], a->s));
}
static void axpy(fe* dst, const fe* src, size_t len, fe s, int nthreads) {
  axpy_arg a = {dst, src, s};
  parallel_for(len, len >= 4096 ? nthreads : 1, axpy_body, &a);
}

/* &poly / (X - point), quotient only (DensePolynomial long division by a monic linear divisor) */
static void divide_linear(const fe* p, size_t len, fe point, fe* q) {
  fe carry = f_zero();
  for (size_t i = len - 1; i >= 1; i--) {
    carry = f_add(&FR, p[i], f_mul(&FR, carry, point));
    q[i - 1] = carry;
  }
}

/* PlonkKzgSnark::preprocess: selector / sigma evaluations -> coefficient polynomials + commitments */
int capcpu_preprocess(unsigned log_n, const u64* sel_evals, const u64* sig_evals, const u64* srs_xy, int nthreads,
                      u64* sel_coef, u64* sig_coef, u64* sel_comms, u64* sig_comms) {
  const size_t n = (size_t)1 << log_n;
  memcpy(sel_coef, sel_evals, 13 * n * 32);
  memcpy(sig_coef, sig_evals, 5 * n * 32);
  for (int s = 0; s < 13; s++) fft_inplace((fe*)sel_coef + s * n, log_n, 1, 0, nthreads);
  for (int s = 0; s < 5; s++) fft_inplace((fe*)sig_coef + s * n, log_n, 1, 0, nthreads);
  for (int s = 0; s < 13; s++) { g1a c = kzg_commit((const g1a*)srs_xy, (fe*)sel_coef + s * n, n, nthreads); memcpy(sel_comms + 8 * s, &c, 64); }
  for (int s = 0; s < 5; s++) { g1a c = kzg_commit((const g1a*)srs_xy, (fe*)sig_coef + s * n, n, nthreads); memcpy(sig_comms + 8 * s, &c, 64); }
  return 0;
}

/* Rounds 1-5 of the prover for one circuit.  All field inputs Montgomery, ABI layout of include/capgpu.h.
 * sig_evals: sigma_i(omega^j) (what jf-relation reads through extended_id_permutation[wire_permutation]). */
int capcpu_prove(unsigned log_n, size_t num_inputs, const u64* sel_coef_in, const u64* sig_coef_in, const u64* sig_evals_in,
                 const u64* k_in, const u64* srs_xy, const u64* sel_comms, const u64* sig_comms, const u64* wires_in,
                 const u64* pub_in, const u64* blinders_in, const uint8_t* ext_msg, size_t ext_len, int nthreads, proof_t* out) {
  const field_t* F = &FR;
  const size_t n = (size_t)1 << log_n, m = 8 * n;
  const unsigned log_m = log_n + 3;
  const g1a* srs = (const g1a*)srs_xy;
  const fe* k = (const fe*)k_in;
  const fe* bl = (const fe*)blinders_in;
  const fe* sel_coef = (const fe*)sel_coef_in;
  const fe* sig_coef = (const fe*)sig_coef_in;
  const fe* sig_ev = (const fe*)sig_evals_in;
  const fe* wires = (const fe*)wires_in;
  const fe one = f_one(F);
  fe omega = omega_of(log_n);
  int rc = 0;

  transcript_t tr; tr_init(&tr);
  if (ext_len) tr_append(&tr, ext_msg, ext_len);
  { u64 v = 254; tr_append(&tr, &v, 8); v = n; tr_append(&tr, &v, 8); v = num_inputs; tr_append(&tr, &v, 8); }
  for (int i = 0; i < 5; i++) tr_append_fr(&tr, k[i]);
  for (int i = 0; i < 13; i++) tr_append_g1(&tr, (const g1a*)(sel_comms + 8 * i));
  for (int i = 0; i < 5; i++) tr_append_g1(&tr, (const g1a*)(sig_comms + 8 * i));
  for (size_t i = 0; i < num_inputs; i++) tr_append_fr(&tr, ((const fe*)pub_in)[i]);

  /* Round 1: wire polynomials (ifft + mask), commitments, public-input polynomial */
  fe* wp[5];
  for (int i = 0; i < 5; i++) {
    wp[i] = (fe*)calloc(m, sizeof(fe));
    memcpy(wp[i], wires + i * n, n * sizeof(fe));
    fft_inplace(wp[i], log_n, 1, 0, nthreads);
    for (int t = 0; t < 2; t++) { wp[i][t] = f_sub(F, wp[i][t], bl[2 * i + t]); wp[i][n + t] = bl[2 * i + t]; }
    g1a c = kzg_commit(srs, wp[i], n + 2, nthreads);
    memcpy(out->wires_poly_comms[i], &c, 64);
  }
  fe* pi = (fe*)calloc(m, sizeof(fe));
  memcpy(pi, pub_in, num_inputs * sizeof(fe));
  fft_inplace(pi, log_n, 1, 0, nthreads);
  for (int i = 0; i < 5; i++) tr_append_g1(&tr, (const g1a*)out->wires_poly_comms[i]);

  /* Round 2: permutation grand product (serial, one division per row) */
  fe beta = tr_challenge(&tr), gamma = tr_challenge(&tr);
  fe* z = (fe*)calloc(m, sizeof(fe));
  {
    fe wj = one;
    z[0] = one;
    for (size_t j = 0; j + 1 < n; j++) {
      fe a = one, b = one;
      for (int i = 0; i < 5; i++) {
        fe tmp = f_add(F, wires[i * n + j], gamma);
        a = f_mul(F, a, f_add(F, tmp, f_mul(F, beta, f_mul(F, k[i], wj))));
        b = f_mul(F, b, f_add(F, tmp, f_mul(F, beta, sig_ev[i * n + j])));
      }
      z[j + 1] = f_mul(F, f_mul(F, z[j], a), f_inv(F, b));
      wj = f_mul(F, wj, omega);
    }
  }
  fft_inplace(z, log_n, 1, 0, nthreads);
  for (int t = 0; t < 3; t++) { z[t] = f_sub(F, z[t], bl[10 + t]); z[n + t] = bl[10 + t]; }
  { g1a c = kzg_commit(srs, z, n + 3, nthreads); memcpy(out->prod_perm_poly_comm, &c, 64); }
  tr_append_g1(&tr, (const g1a*)out->prod_perm_poly_comm);

  /* Round 3: quotient polynomial -- 25 coset FFTs of size 8n (par_iter over polynomials), point-wise
   * evaluation (par_iter over points), one coset IFFT, split + mask, 5 commitments */
  fe alpha = tr_challenge(&tr);
  fe* cos[25];
  for (int s = 0; s < 13; s++) { cos[s] = (fe*)calloc(m, sizeof(fe)); memcpy(cos[s], sel_coef + s * n, n * sizeof(fe)); }
  for (int s = 0; s < 5; s++) { cos[13 + s] = (fe*)calloc(m, sizeof(fe)); memcpy(cos[13 + s], sig_coef + s * n, n * sizeof(fe)); }
  for (int s = 0; s < 5; s++) { cos[18 + s] = (fe*)malloc(m * sizeof(fe)); memcpy(cos[18 + s], wp[s], m * sizeof(fe)); }
  cos[23] = (fe*)malloc(m * sizeof(fe)); memcpy(cos[23], z, m * sizeof(fe));
  cos[24] = (fe*)malloc(m * sizeof(fe)); memcpy(cos[24], pi, m * sizeof(fe));
  {
    cfft_arg ca = {cos, NULL, log_m, 1};
    if (nthreads >= 8) parallel_for(25, nthreads, coset_fft_many, &ca);
    else { ca.nthreads_inner = nthreads; coset_fft_many(0, 25, &ca); }
  }
  fe* t = (fe*)malloc(m * sizeof(fe));
  {
    quot_arg q;
    q.n = n; q.m = m; q.k = k; q.sel = cos; q.sig = cos + 13; q.w = cos + 18; q.z = cos[23]; q.pi = cos[24]; q.out = t;
    q.alpha = alpha; q.alpha2 = f_sqr(F, alpha); q.beta = beta; q.gamma = gamma; q.wm = omega_of(log_m);
    fe g; memcpy(g.v, GEN5, 32);
    fe x = g;
    for (int i = 0; i < 8; i++) { q.zh_inv[i] = f_inv(F, f_sub(F, f_pow_u64(F, x, n), one)); x = f_mul(F, x, q.wm); }
    parallel_for(m, nthreads, quot_body, &q);
  }
  fft_inplace(t, log_m, 1, 1, nthreads);
  const size_t deg = 5 * n + 7;
  for (size_t i = deg + 1; i < m; i++) if (!f_is_zero(t[i])) rc = -3;
  if (f_is_zero(t[deg])) rc = -3;
  fe* sp[5];
  size_t sp_len[5];
  for (int i = 0; i < 5; i++) {
    sp[i] = (fe*)calloc(n + 3, sizeof(fe));
    size_t lo = i * (n + 2), hi = i < 4 ? (i + 1) * (n + 2) : deg + 1;
    memcpy(sp[i], t + lo, (hi - lo) * sizeof(fe));
    sp_len[i] = hi - lo;
  }
  {
    fe last = f_zero();
    for (int i = 0; i < 4; i++) { sp[i][0] = f_sub(F, sp[i][0], last); sp[i][n + 2] = bl[13 + i]; sp_len[i] = n + 3; last = bl[13 + i]; }
    sp[4][0] = f_sub(F, sp[4][0], last);
  }
  for (int i = 0; i < 5; i++) { g1a c = kzg_commit(srs, sp[i], sp_len[i], nthreads); memcpy(out->split_quot_poly_comms[i], &c, 64); }
  for (int i = 0; i < 5; i++) tr_append_g1(&tr, (const g1a*)out->split_quot_poly_comms[i]);

  /* Round 4: evaluations */
  fe zeta = tr_challenge(&tr);
  fe zeta_w = f_mul(F, zeta, omega);
  fe ev[10];
  {
    fe* polys[10]; size_t lens[10]; fe xs[10];
    for (int i = 0; i < 5; i++) { polys[i] = wp[i]; lens[i] = n + 2; xs[i] = zeta; }
    for (int i = 0; i < 4; i++) { polys[5 + i] = (fe*)(sig_coef + i * n); lens[5 + i] = n; xs[5 + i] = zeta; }
    polys[9] = z; lens[9] = n + 3; xs[9] = zeta_w;
    eval_arg ea = {polys, lens, xs, ev};
    parallel_for(10, nthreads, eval_body, &ea);
  }
  memcpy(out->wires_evals, ev, 5 * 32); memcpy(out->wire_sigma_evals, ev + 5, 4 * 32); memcpy(out->perm_next_eval, ev + 9, 32);
  for (int i = 0; i < 10; i++) tr_append_fr(&tr, ev[i]);

  /* linearisation polynomial */
  fe* lin = (fe*)calloc(n + 3, sizeof(fe));
  {
    fe* w = ev; fe* se = ev + 5; fe zw = ev[9];
    fe zh = f_sub(F, f_pow_u64(F, zeta, n), one);
    fe l1 = f_mul(F, zh, f_inv(F, f_mul(F, f_from_u64(F, n), f_sub(F, zeta, one))));
    fe w01 = f_mul(F, w[0], w[1]), w23 = f_mul(F, w[2], w[3]);
    fe sc[13];
    for (int i = 0; i < 4; i++) sc[i] = w[i];
    sc[4] = w01; sc[5] = w23;
    for (int i = 0; i < 4; i++) { fe w2 = f_sqr(F, w[i]); sc[6 + i] = f_mul(F, f_sqr(F, w2), w[i]); }
    sc[10] = f_neg(F, w[4]); sc[11] = one; sc[12] = f_mul(F, f_mul(F, w01, w23), w[4]);
    for (int s = 0; s < 13; s++) axpy(lin, sel_coef + s * n, n, sc[s], nthreads);
    fe cz = alpha, bz = f_mul(F, beta, zeta);
    for (int j = 0; j < 5; j++) cz = f_mul(F, cz, f_add(F, f_add(F, w[j], f_mul(F, k[j], bz)), gamma));
    cz = f_add(F, cz, f_mul(F, f_sqr(F, alpha), l1));
    axpy(lin, z, n + 3, cz, nthreads);
    fe cs = f_mul(F, f_mul(F, alpha, beta), zw);
    for (int j = 0; j < 4; j++) cs = f_mul(F, cs, f_add(F, f_add(F, w[j], f_mul(F, beta, se[j])), gamma));
    axpy(lin, sig_coef + 4 * n, n, f_neg(F, cs), nthreads);
    fe zn2 = f_mul(F, f_mul(F, f_add(F, zh, one), zeta), zeta), c = one;
    for (int i = 0; i < 5; i++) { axpy(lin, sp[i], sp_len[i], f_neg(F, f_mul(F, zh, c)), nthreads); c = f_mul(F, c, zn2); }
  }

  /* Round 5: batched opening proofs */
  fe v = tr_challenge(&tr);
  fe* batch = (fe*)calloc(n + 3, sizeof(fe));
  memcpy(batch, lin, (n + 3) * sizeof(fe));
  {
    fe c = v;
    for (int i = 0; i < 5; i++) { axpy(batch, wp[i], n + 2, c, nthreads); c = f_mul(F, c, v); }
    for (int i = 0; i < 4; i++) { axpy(batch, sig_coef + i * n, n, c, nthreads); c = f_mul(F, c, v); }
  }
  fe* q1 = (fe*)calloc(n + 3, sizeof(fe));
  fe* q2 = (fe*)calloc(n + 3, sizeof(fe));
  divide_linear(batch, n + 3, zeta, q1);
  divide_linear(z, n + 3, zeta_w, q2);
  { g1a c = kzg_commit(srs, q1, n + 2, nthreads); memcpy(out->opening_proof, &c, 64); }
  { g1a c = kzg_commit(srs, q2, n + 2, nthreads); memcpy(out->shifted_opening_proof, &c, 64); }

  for (int i = 0; i < 5; i++) { free(wp[i]); free(sp[i]); }
  for (int i = 0; i < 25; i++) free(cos[i]);
  free(pi); free(z); free(t); free(lin); free(batch); free(q1); free(q2); free(tr.buf);
  return rc;
}

/* standalone primitives for the kernel sweeps (BASELINE config 4) */
int capcpu_msm(const u64* srs_xy, const u64* scalars, size_t n, int scalars_mont, int nthreads, u64* out_xy) {
  g1a r;
  if (scalars_mont) r = kzg_commit((const g1a*)srs_xy, (const fe*)scalars, n, nthreads);
  else r = msm_arkworks((const g1a*)srs_xy, (const fe*)scalars, n, nthreads);
  memcpy(out_xy, &r, 64);
  return 0;
}

int capcpu_ntt(u64* data, unsigned log_n, int inverse, int coset, int nthreads) {
  fft_inplace((fe*)data, log_n, inverse, coset, nthreads);
  return 0;
}

/* synthetic SRS tau^i * G (KZG10::setup shape): fixed-base 8-bit window table (32 x 255 affine
 * multiples of G), then <= 32 mixed additions per power, powers spread over the threads */
typedef struct { const fe* pows; const g1a* table; u64* out; } srs_arg;
static void srs_body(size_t lo, size_t hi, void* argp) {
  srs_arg* a = (srs_arg*)argp;
  for (size_t i = lo; i < hi; i++) {
    fe e = f_from_mont(&FR, a->pows[i]);
    g1j acc = g1j_inf();
    for (int w = 0; w < 32; w++) {
      unsigned d = (unsigned)((e.v[w >> 3] >> ((w & 7) * 8)) & 0xff);
      if (d) g1j_add_mixed(&acc, &a->table[w * 256 + d]);
    }
    g1a r = g1j_to_affine(&acc);
    memcpy(a->out + 8 * i, &r, 64);
  }
}
int capcpu_srs(const u64* tau_mont, size_t n, int nthreads, u64* out_xy) {
  fe tau; memcpy(tau.v, tau_mont, 32);
  fe* pows = (fe*)malloc((n ? n : 1) * sizeof(fe));
  fe s = f_one(&FR);
  for (size_t i = 0; i < n; i++) { pows[i] = s; s = f_mul(&FR, s, tau); }
  g1a* table = (g1a*)calloc(32 * 256, sizeof(g1a));
  g1a g; g.x = f_one(&FQ); g.y = f_dbl(&FQ, f_one(&FQ));
  g1j base; base.x = g.x; base.y = g.y; base.z = f_one(&FQ);
  for (int w = 0; w < 32; w++) {
    g1a ba = g1j_to_affine(&base);
    g1j cur = g1j_inf();
    for (int d = 1; d < 256; d++) { g1j_add_mixed(&cur, &ba); table[w * 256 + d] = g1j_to_affine(&cur); }
    for (int k = 0; k < 8; k++) g1j_double(&base);
  }
  srs_arg a = {pows, table, out_xy};
  parallel_for(n, nthreads, srs_body, &a);
  free(pows); free(table);
  return 0;
}
