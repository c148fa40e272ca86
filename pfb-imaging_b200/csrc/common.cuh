// Shared device helpers of the B200 w-gridder kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PFBG_TILE 16
#define PFBG_MAXW 16

template <typename T> struct cplx_of;
template <> struct cplx_of<float> { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };

// Geometry + kernel parameters, passed by value to every kernel.
struct GParams {
  int nx, ny, nu, nv, W, nplanes, nchan;
  int do_wgridding, divide_by_n;
  double beta, pixsize_x, pixsize_y, center_x, center_y;
  double usign, vsign, wsign, w0, dw, nshift;
  int ntile_u, ntile_v;
  // Mirror planes (pmirror > 0): plane p sits at w_p = (p + 1/2) dw, so plane -p-1 is the Hermitian mirror of
  // plane p (real image): G_{-p-1}(u,v) = conj G_p(-u,-v).  Samples near w = 0 whose support reaches planes
  // -pmirror..-1 read / write them at the mirrored cell of plane -q-1, conjugated; those planes are never
  // stored or transformed.
  int pmirror;
  int fast_screen;  // fp32: w-screen phasors from the SFU (epsilon >= 3e-6)
  // Batched snapshots (pfb hci: thousands of small images of one geometry, utils/stokes2im.py:635-683): nbatch
  // images share this plan; snapshot s owns rows [row offsets], planes [snap_pbase[s], + snap_np[s]) of ONE plane
  // stack of `nplanes` planes in total, its plane 0 sits at w = snap_w0[s].  All null / 1 for an ordinary plan.
  int nbatch;
  const int* row_snap;      // (nrow) snapshot of every row
  const double* snap_w0;    // (nbatch)
  const int* snap_pbase;    // (nbatch)
  const int* snap_np;       // (nbatch)
};

// ---------------------------------------------------------------------------
// Bit-exact coordinate arithmetic (Appendix B of SURVEY.md / oracle.bin_indices):
// every operation is one correctly rounded IEEE fp64 op, no FMA contraction.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void axis_coord(double t, double pix, int n, int W, double& g, int& i0) {
  double x = __dmul_rn(t, pix);
  double f = __dsub_rn(x, floor(x));
  g = __dadd_rn(__dmul_rn(f, (double)n), 0.5 * (double)n);
  if (g >= (double)n) g = __dsub_rn(g, (double)n);
  i0 = (int)floor(__dsub_rn(g, 0.5 * (double)W)) + 1;
}

struct VisCoord {
  double ut, vt, wt;  // wavelengths, signs applied
  double gu, gv, gw;  // grid coordinates
  int iu0, iv0, ip0;  // first touched cell / plane
  int conj;           // sample folded onto w >= 0: (u,v,w) -> -(u,v,w), visibility conjugated
};

__device__ __forceinline__ VisCoord vis_coord(const GParams& p, const double* __restrict__ uvw,
                                              const double* __restrict__ fscale, int64_t row, int chan) {
  VisCoord c;
  double s = fscale[chan];
  c.ut = __dmul_rn(__dmul_rn(p.usign, uvw[3 * row + 0]), s);
  c.vt = __dmul_rn(__dmul_rn(p.vsign, uvw[3 * row + 1]), s);
  c.wt = __dmul_rn(__dmul_rn(p.wsign, uvw[3 * row + 2]), s);
  // Hermitian fold (the image is real, so V(-u,-v,-w) = conj V(u,v,w)): samples with w < 0 are
  // gridded at -(u,v,w) with the conjugate visibility, which halves the w-range the planes cover.
  c.conj = 0;
  if (p.do_wgridding && c.wt < 0.0) { c.ut = -c.ut; c.vt = -c.vt; c.wt = -c.wt; c.conj = 1; }
  axis_coord(c.ut, p.pixsize_x, p.nu, p.W, c.gu, c.iu0);
  axis_coord(c.vt, p.pixsize_y, p.nv, p.W, c.gv, c.iv0);
  double w0 = p.w0;
  int pbase = 0, np_ = p.nplanes;
  if (p.row_snap) {  // batched snapshots: this row's own plane block
    const int sn = p.row_snap[row];
    w0 = p.snap_w0[sn]; pbase = p.snap_pbase[sn]; np_ = p.snap_np[sn];
  }
  if (p.do_wgridding) {
    c.gw = __ddiv_rn(__dsub_rn(c.wt, w0), p.dw);
    int ip = (int)floor(__dsub_rn(c.gw, 0.5 * (double)p.W)) + 1;
    int lo = -p.pmirror, hi = np_ - p.W;
    c.ip0 = ip < lo ? lo : (ip > hi ? hi : ip);
  } else {
    c.gw = 0.0;
    c.ip0 = 0;
  }
  if (pbase) { c.gw += (double)pbase; c.ip0 += pbase; }  // absolute plane units (exact: small integers)
  return c;
}

__device__ __forceinline__ int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

__device__ __forceinline__ uint64_t bucket_key(const GParams& p, const VisCoord& c) {
  int iuw = wrap(c.iu0, p.nu), ivw = wrap(c.iv0, p.nv);
  uint64_t tile = (uint64_t)(iuw / PFBG_TILE) * (uint64_t)p.ntile_v + (uint64_t)(ivw / PFBG_TILE);
  uint64_t fine = (uint64_t)((iuw % PFBG_TILE) * PFBG_TILE + (ivw % PFBG_TILE));
  return (tile * (uint64_t)(p.nplanes + p.pmirror) + (uint64_t)(c.ip0 + p.pmirror)) * (uint64_t)(PFBG_TILE * PFBG_TILE) + fine;
}

// element offset of cell (iu, iv) of plane pl in the stack; planes pl < 0 live at the mirrored cell of plane
// -pl-1 and are conjugated (cj)
__device__ __forceinline__ int64_t plane_cell(const GParams& p, int pl, int iu, int iv, bool& cj) {
  const int64_t plane_sz = (int64_t)p.nu * p.nv;
  if (pl >= 0) { cj = false; return (int64_t)pl * plane_sz + (int64_t)iu * p.nv + iv; }
  cj = true;
  const int mu = iu ? p.nu - iu : 0, mv = iv ? p.nv - iv : 0;
  return (int64_t)(-pl - 1) * plane_sz + (int64_t)mu * p.nv + mv;
}

// ES kernel phi(x) = exp(beta (sqrt(1-x^2) - 1)), zero outside |x| <= 1.
__device__ __forceinline__ float es_eval(float x, float beta) {
  float a = (1.0f - x) * (1.0f + x);
  return a < 0.0f ? 0.0f : expf(beta * (sqrtf(a) - 1.0f));
}
__device__ __forceinline__ double es_eval(double x, double beta) {
  double a = (1.0 - x) * (1.0 + x);
  return a < 0.0 ? 0.0 : exp(beta * (sqrt(a) - 1.0));
}

// weight of cell (i0 + j) for coordinate g
template <typename T>
__device__ __forceinline__ T tap(double g, int i0, int j, int W, T beta) {
  double x = ((double)(i0 + j) - g) * (2.0 / (double)W);
  return es_eval((T)x, beta);
}

// e^{2 pi i t} with the fp64 phase reduced to [-1/2, 1/2] turns first
__device__ __forceinline__ void cis_turns(double t, float& c, float& s) {
  t -= rint(t);
  sincospif(2.0f * (float)t, &s, &c);
}
__device__ __forceinline__ void cis_turns(double t, double& c, double& s) {
  t -= rint(t);
  sincospi(2.0 * t, &s, &c);
}

// w-screen phasor: FAST (fp32 only) takes sin / cos from the SFU after the same fp64 range reduction —
// |2 pi t| <= pi, where sin.approx / cos.approx are good to ~4e-7 absolute — 5 instructions instead of ~30
template <bool FAST, typename T>
__device__ __forceinline__ void cis_screen(double t, T& c, T& s) {
  if constexpr (FAST && sizeof(T) == 4) {
    t -= rint(t);
    const float a = 6.283185307179586f * (float)t;
    s = __sinf(a);
    c = __cosf(a);
  } else {
    cis_turns(t, c, s);
  }
}

// phase (turns) of the centre shift and the n-1 shift for one sample
__device__ __forceinline__ double vis_phase_turns(const GParams& p, const VisCoord& c) {
  double t = c.ut * p.center_x + c.vt * p.center_y;
  if (p.do_wgridding) t += c.wt * p.nshift;
  return t;
}

// n - 1 for pixel (i, j), numerically stable
__device__ __forceinline__ double pixel_nm1(const GParams& p, int i, int j) {
  double l = p.center_x + (double)(i - p.nx / 2) * p.pixsize_x;
  double m = p.center_y + (double)(j - p.ny / 2) * p.pixsize_y;
  double r2 = l * l + m * m;
  return -r2 / (sqrt(1.0 - r2) + 1.0);
}

// Packed complex helpers.  sm_100 has two-wide fp32 FMA/MUL (FFMA2 / FMUL2, scalar second operand
// broadcast): same flop rate as FFMA but half the issue slots, which is what bounds the run kernels.
__device__ __forceinline__ float2 cfma_s(float2 a, float s, float2 c) {  // a * s + c
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 cmul_s(float2 a, float s) {  // a * s
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(s));
  return d;
}
__device__ __forceinline__ double2 cfma_s(double2 a, double s, double2 c) { return make_double2(a.x * s + c.x, a.y * s + c.y); }
__device__ __forceinline__ double2 cmul_s(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

__device__ __forceinline__ void atomic_add_c(float2* a, float re, float im) {
  atomicAdd(a, make_float2(re, im));  // sm_90+: one vector RED
}
__device__ __forceinline__ void atomic_add_c(double2* a, double re, double im) {
  atomicAdd(&a->x, re);
  atomicAdd(&a->y, im);
}
