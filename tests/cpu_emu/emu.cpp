// Host emulation harness: runs the EXACT device algorithms of cap_b200/csrc/{fp,ec}.cuh on
// the CPU (PTX carry semantics emulated) so they can be checked against the big-int oracle
// without a GPU.  Test infrastructure only -- never linked into the product library.
#define CAPGPU_HOST_EMU 1
#include "../../cap_b200/csrc/fp.cuh"
#include <string.h>
using namespace capgpu;

extern "C" {
void emu_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fr z = fp_mul(x, y); memcpy(r, z.v, 32); }
void emu_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fq z = fp_mul(x, y); memcpy(r, z.v, 32); }
void emu_fq_mul_add(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  Fq x, y, z, w; memcpy(x.v, a, 32); memcpy(y.v, b, 32); memcpy(z.v, c, 32); memcpy(w.v, d, 32);
  Fq o = fp_mul_add(x, y, z, w); memcpy(r, o.v, 32);
}
void emu_fr_mul_sub(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  Fr x, y, z, w; memcpy(x.v, a, 32); memcpy(y.v, b, 32); memcpy(z.v, c, 32); memcpy(w.v, d, 32);
  Fr o = fp_mul_sub(x, y, z, w); memcpy(r, o.v, 32);
}
void emu_fr_sqr(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_sqr(x); memcpy(r, z.v, 32); }
void emu_fq_sqr(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.v, a, 32); Fq z = fp_sqr(x); memcpy(r, z.v, 32); }
void emu_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fr z = fp_add(x, y); memcpy(r, z.v, 32); }
void emu_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fr z = fp_sub(x, y); memcpy(r, z.v, 32); }
void emu_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fq z = fp_add(x, y); memcpy(r, z.v, 32); }
void emu_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fq x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); Fq z = fp_sub(x, y); memcpy(r, z.v, 32); }
void emu_fr_neg(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_neg(x); memcpy(r, z.v, 32); }
void emu_fr_inv(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_inv(x); memcpy(r, z.v, 32); }
void emu_fq_inv(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.v, a, 32); Fq z = fp_inv(x); memcpy(r, z.v, 32); }
void emu_fr_from_mont(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_from_mont(x); memcpy(r, z.v, 32); }
void emu_fr_to_mont(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_to_mont(x); memcpy(r, z.v, 32); }
}

#include "../../cap_b200/csrc/ec.cuh"
extern "C" {
// points: affine 16 x u32 (x||y Montgomery, zeros = infinity)
void emu_g1_add_mixed_chain(const uint32_t* pts, const int* negs, int n, uint32_t* out) {
  G1XYZZ acc = G1XYZZ::inf();
  for (int i = 0; i < n; i++) {
    G1Affine p; memcpy(&p, pts + 16 * i, 64);
    if (!p.is_inf()) xyzz_add_mixed(acc, p.x, p.y, negs[i] != 0);
  }
  G1Affine r = xyzz_to_affine(acc); memcpy(out, &r, 64);
}
// the bucket walk of msm_accumulate (msm.cu): the first two entries through the affine + affine formula when both are finite
// and their x differ, everything else through the mixed addition
void emu_g1_bucket_chain(const uint32_t* pts, const int* negs, int n, uint32_t* out) {
  G1XYZZ acc = G1XYZZ::inf();
  int e = 0;
  if (n >= 2) {
    G1Affine p0, p1; memcpy(&p0, pts, 64); memcpy(&p1, pts + 16, 64);
    if (!p0.is_inf() && !p1.is_inf()) {
      if (xyzz_set_affine2(acc, p0.x, negs[0] ? fp_neg(p0.y) : p0.y, p1.x, negs[1] ? fp_neg(p1.y) : p1.y)) e = 2;
    }
  }
  for (; e < n; e++) {
    G1Affine p; memcpy(&p, pts + 16 * e, 64);
    if (!p.is_inf()) xyzz_add_mixed(acc, p.x, p.y, negs[e] != 0);
  }
  G1Affine r = xyzz_to_affine(acc); memcpy(out, &r, 64);
}
void emu_g1_add_full(const uint32_t* pts_a, int na, const uint32_t* pts_b, int nb, uint32_t* out) {
  G1XYZZ a = G1XYZZ::inf(), b = G1XYZZ::inf();
  for (int i = 0; i < na; i++) { G1Affine p; memcpy(&p, pts_a + 16 * i, 64); if (!p.is_inf()) xyzz_add_mixed(a, p.x, p.y, false); }
  for (int i = 0; i < nb; i++) { G1Affine p; memcpy(&p, pts_b + 16 * i, 64); if (!p.is_inf()) xyzz_add_mixed(b, p.x, p.y, false); }
  xyzz_add(a, b);
  G1Affine r = xyzz_to_affine(a); memcpy(out, &r, 64);
}
void emu_g1_mul_small(const uint32_t* pt, uint32_t k, uint32_t* out) {
  G1Affine p; memcpy(&p, pt, 64);
  G1XYZZ r = xyzz_mul_small(xyzz_from_affine(p), k);
  G1Affine a = xyzz_to_affine(r); memcpy(out, &a, 64);
}
}

extern "C" {
void emu_fr_inv_fermat(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_inv_fermat(x); memcpy(r, z.v, 32); }
void emu_fq_inv_fermat(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.v, a, 32); Fq z = fp_inv_fermat(x); memcpy(r, z.v, 32); }
void emu_fr_inv_euclid(const uint32_t* a, uint32_t* r) { Fr x; memcpy(x.v, a, 32); Fr z = fp_inv_euclid(x); memcpy(r, z.v, 32); }
void emu_fq_inv_euclid(const uint32_t* a, uint32_t* r) { Fq x; memcpy(x.v, a, 32); Fq z = fp_inv_euclid(x); memcpy(r, z.v, 32); }
}


// ---- 12-limb fields / G1 of BLS12-381 (curve 1) and BLS12-377 (curve 2): fpn.cuh -----------------------
#include "../../cap_b200/csrc/fpn.cuh"
template <class F>
static void fqn_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  F x, y, z;
  memcpy(x.v, a, sizeof x.v);
  if (b) memcpy(y.v, b, sizeof y.v); else y = F::zero();
  switch (op) {
    case 0: z = fp_mul(x, y); break;
    case 1: z = fp_sqr(x); break;
    case 2: z = fp_inv(x); break;
    case 3: z = fp_add(x, y); break;
    case 4: z = fp_sub(x, y); break;
    case 5: z = fp_inv_fermat(x); break;
    case 6: z = fp_neg(x); break;
    default: z = fp_from_mont(x); break;
  }
  memcpy(r, z.v, sizeof z.v);
}
// sum of +-points by mixed additions, then + (sum of a second list) by a full addition, doubled `dbl` times, to affine
template <class F>
static void g1n_chain(const uint32_t* pts_a, const int* negs, int na, const uint32_t* pts_b, int nb, int dbl, uint32_t* out) {
  G1XyzzT<F> a = G1XyzzT<F>::inf(), b = G1XyzzT<F>::inf();
  for (int i = 0; i < na; i++) { G1AffineT<F> p; memcpy(&p, pts_a + 2 * F::N * i, sizeof p); if (!p.is_inf()) xyzz_add_mixed(a, p.x, p.y, negs[i] != 0); }
  for (int i = 0; i < nb; i++) { G1AffineT<F> p; memcpy(&p, pts_b + 2 * F::N * i, sizeof p); if (!p.is_inf()) xyzz_add_mixed(b, p.x, p.y, false); }
  xyzz_add(a, b);
  for (int i = 0; i < dbl; i++) a = xyzz_dbl(a);
  G1AffineT<F> r = xyzz_to_affine(a);
  memcpy(out, &r, sizeof r);
}
extern "C" {
void emu_fqn_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  if (curve == 1) fqn_op<Fq381>(op, a, b, r); else fqn_op<Fq377>(op, a, b, r);
}
void emu_g1n_chain(int curve, const uint32_t* pts_a, const int* negs, int na, const uint32_t* pts_b, int nb, int dbl, uint32_t* out) {
  if (curve == 1) g1n_chain<Fq381>(pts_a, negs, na, pts_b, nb, dbl, out); else g1n_chain<Fq377>(pts_a, negs, na, pts_b, nb, dbl, out);
}
}
