// In-shared-memory mixed-radix FFT engine (radices 16, 8, 4, 2, 3, 5, 7) used by the fused,
// pruned plane transforms in fused_fft.cuh.
//
// Data layout in shared memory: element n of interleaved transform c sits at s[n*C + c]
// (C = 1 for row transforms, C = 4 (fp32) / 2 (fp64) columns for the strided column transforms so
// that every global access is a full 32-byte sector).
//
// Only the FORWARD transform (e^{-2 pi i nk/N}) is implemented; the inverse is conj-in/conj-out.
//   dif: natural-order input, output left in mixed-radix digit-reversed order (pos -> k = rev[pos])
//   dit: digit-reversed input (x[n] stored at posmap[n]), natural-order output
// Both walk the same stage list (L_s, r_s): stage s works on blocks of length L_s = N/(r_1..r_{s-1})
// with butterflies over elements  base + m*(L_s/r_s) + k,  m < r_s.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FFT_MAX_STAGES 12

struct FftDesc {
  int n;
  int nstage;
  int radix[FFT_MAX_STAGES];
};

template <typename T> struct cx2 { T x, y; };

// Shared-memory index padding: one spare element after every 128 bytes, so that the power-of-two
// strides of the late stages (and the digit-reversed scatter) spread over all 32 banks.
template <typename T> __device__ __host__ __forceinline__ int fft_pad(int e) {
  // +1 element per 128 B: spreads the small power-of-two strides of the late stages.  (A deeper
  // skew that also spreads the digit-reversed scatter was measured 20 % SLOWER on B200: these
  // kernels are issue-bound and the extra index arithmetic costs more than the conflicts.)
#ifdef FFT_PAD_V2
  return e + (e >> (sizeof(T) == 4 ? 4 : 3)) + (e >> 7) + (e >> 11);
#else
  return e + (e >> (sizeof(T) == 4 ? 4 : 3));
#endif
}
template <typename T> __device__ __host__ __forceinline__ size_t fft_smem_bytes(int nelem) {
  return (size_t)(fft_pad<T>(nelem - 1) + 1) * sizeof(cx2<T>);
}
template <typename T> __device__ __forceinline__ cx2<T> cmul(cx2<T> a, cx2<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T> __device__ __forceinline__ cx2<T> cadd(cx2<T> a, cx2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ cx2<T> csub(cx2<T> a, cx2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ cx2<T> mul_mi(cx2<T> a) { return {a.y, -a.x}; }  // a * (-i)

template <typename T> __device__ __forceinline__ void bf2(cx2<T>& a, cx2<T>& b) {
  cx2<T> t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
// forward radix-4 on (a,b,c,d) in natural order in/out
template <typename T> __device__ __forceinline__ void bf4(cx2<T>& a, cx2<T>& b, cx2<T>& c, cx2<T>& d) {
  cx2<T> t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = mul_mi(csub(b, d));
  a = cadd(t0, t2);
  b = cadd(t1, t3);
  c = csub(t0, t2);
  d = csub(t1, t3);
}

template <typename T, int R> struct SmallDft;

template <typename T> struct SmallDft<T, 2> {
  __device__ static __forceinline__ void run(cx2<T>* x) { bf2(x[0], x[1]); }
};
template <typename T> struct SmallDft<T, 4> {
  __device__ static __forceinline__ void run(cx2<T>* x) { bf4(x[0], x[1], x[2], x[3]); }
};
template <typename T> struct SmallDft<T, 8> {
  __device__ static __forceinline__ void run(cx2<T>* x) {
    // n = 2 n1 + n2 ; k = k1 + 4 k2 : radix-4 over n1 for each n2, twiddle W8^{n2 k1}, radix-2 over n2
    cx2<T> e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    cx2<T> o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    bf4(e0, e1, e2, e3);
    bf4(o0, o1, o2, o3);
    const T h = (T)0.70710678118654752440;
    o1 = {(o1.x + o1.y) * h, (o1.y - o1.x) * h};    // * (1 - i)/sqrt2
    o2 = mul_mi(o2);                                 // * (-i)
    o3 = {(o3.y - o3.x) * h, -(o3.x + o3.y) * h};   // * (-1 - i)/sqrt2
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, o1); x[5] = csub(e1, o1);
    x[2] = cadd(e2, o2); x[6] = csub(e2, o2);
    x[3] = cadd(e3, o3); x[7] = csub(e3, o3);
  }
};
template <typename T> struct SmallDft<T, 16> {
  __device__ static __forceinline__ void run(cx2<T>* x) {
    // n = 4 n1 + n2, k = k1 + 4 k2: A[n2][k1] = DFT4_{n1} x[4 n1 + n2]; A *= W16^{n2 k1}; X[k1+4k2] = DFT4_{n2} A
    cx2<T> a[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      a[n2][0] = x[n2]; a[n2][1] = x[4 + n2]; a[n2][2] = x[8 + n2]; a[n2][3] = x[12 + n2];
      bf4(a[n2][0], a[n2][1], a[n2][2], a[n2][3]);
    }
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    // W16^m = cos(pi m/8) - i sin(pi m/8)
    const cx2<T> w1 = {c1, -s1}, w2 = {h, -h}, w3 = {s1, -c1}, w6 = {-h, -h}, w9 = {-c1, s1};
    a[1][1] = cmul(a[1][1], w1); a[1][2] = cmul(a[1][2], w2); a[1][3] = cmul(a[1][3], w3);
    a[2][1] = cmul(a[2][1], w2); a[2][2] = mul_mi(a[2][2]);   a[2][3] = cmul(a[2][3], w6);
    a[3][1] = cmul(a[3][1], w3); a[3][2] = cmul(a[3][2], w6); a[3][3] = cmul(a[3][3], w9);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      bf4(a[0][k1], a[1][k1], a[2][k1], a[3][k1]);
      x[k1] = a[0][k1]; x[k1 + 4] = a[1][k1]; x[k1 + 8] = a[2][k1]; x[k1 + 12] = a[3][k1];
    }
  }
};
template <typename T> struct SmallDft<T, 3> {
  __device__ static __forceinline__ void run(cx2<T>* x) {
    const T s = (T)0.86602540378443864676;
    cx2<T> t = cadd(x[1], x[2]);
    cx2<T> d = csub(x[1], x[2]);
    cx2<T> m = {x[0].x - (T)0.5 * t.x, x[0].y - (T)0.5 * t.y};
    cx2<T> r = {s * d.y, -s * d.x};  // -i s d
    x[0] = cadd(x[0], t);
    x[1] = cadd(m, r);
    x[2] = csub(m, r);
  }
};
// odd primes 5 and 7: direct O(R^2) evaluation with compile-time roots (rare stages)
template <int R> __device__ __forceinline__ void root_of_unity(int m, double& c, double& s);
template <> __device__ __forceinline__ void root_of_unity<5>(int m, double& c, double& s) {
  constexpr double C[5] = {1.0, 0.30901699437494742410, -0.80901699437494742410, -0.80901699437494742410, 0.30901699437494742410};
  constexpr double S[5] = {0.0, 0.95105651629515357212, 0.58778525229247312917, -0.58778525229247312917, -0.95105651629515357212};
  c = C[m]; s = S[m];
}
template <> __device__ __forceinline__ void root_of_unity<7>(int m, double& c, double& s) {
  constexpr double C[7] = {1.0, 0.62348980185873353053, -0.22252093395631440429, -0.90096886790241912624,
                           -0.90096886790241912624, -0.22252093395631440429, 0.62348980185873353053};
  constexpr double S[7] = {0.0, 0.78183148246802980871, 0.97492791218182360702, 0.43388373911755812048,
                           -0.43388373911755812048, -0.97492791218182360702, -0.78183148246802980871};
  c = C[m]; s = S[m];
}
template <> __device__ __forceinline__ void root_of_unity<11>(int m, double& c, double& s) {
  constexpr double C[11] = {1.0, 0.84125353283118120551, 0.41541501300188643508, -0.14231483827328500480,
                            -0.65486073394528498959, -0.95949297361449736865, -0.95949297361449736865,
                            -0.65486073394528498959, -0.14231483827328500480, 0.41541501300188643508,
                            0.84125353283118120551};
  constexpr double S[11] = {0.0, 0.54064081745559755543, 0.90963199535451833011, 0.98982144188093279524,
                            0.75574957435425826890, 0.28173255684142967104, -0.28173255684142967104,
                            -0.75574957435425826890, -0.98982144188093279524, -0.90963199535451833011,
                            -0.54064081745559755543};
  c = C[m]; s = S[m];
}
template <typename T, int R> __device__ __forceinline__ void dft_direct(cx2<T>* x) {
  cx2<T> y[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    cx2<T> acc = x[0];
#pragma unroll
    for (int n = 1; n < R; ++n) {
      double c, s;
      root_of_unity<R>((n * k) % R, c, s);  // e^{-2 pi i nk/R} = c - i s
      const T cc = (T)c, ss = (T)s;
      acc.x += x[n].x * cc + x[n].y * ss;
      acc.y += x[n].y * cc - x[n].x * ss;
    }
    y[k] = acc;
  }
#pragma unroll
  for (int k = 0; k < R; ++k) x[k] = y[k];
}
template <typename T> struct SmallDft<T, 5> {
  __device__ static __forceinline__ void run(cx2<T>* x) { dft_direct<T, 5>(x); }
};
template <typename T> struct SmallDft<T, 7> {
  __device__ static __forceinline__ void run(cx2<T>* x) { dft_direct<T, 7>(x); }
};
template <typename T> struct SmallDft<T, 11> {  // ducc0.fft.good_size admits the factor 11 (PSF grids)
  __device__ static __forceinline__ void run(cx2<T>* x) { dft_direct<T, 11>(x); }
};

// ---------------------------------------------------------------------------
// one stage over the whole shared-memory array (all threads of the CTA), forward transform
//   DIF: butterfly, then output m *= W_L^{k m};   DIT: input m *= W_L^{k m}, then butterfly
// tw[t] = e^{-2 pi i t/N} (global memory, L1/L2 resident), tstride = N / L.
// ---------------------------------------------------------------------------
// w^m from the binary powers w1, w2, w4, w8 of ONE table load (product depth <= 3, few registers)
template <typename T, int M_>
__device__ __forceinline__ cx2<T> twiddle_pow(const cx2<T>& w1, const cx2<T>& w2, const cx2<T>& w4, const cx2<T>& w8) {
  cx2<T> r = {(T)1, (T)0};
  bool have = false;
  if (M_ & 8) { r = w8; have = true; }
  if (M_ & 4) { r = have ? cmul(r, w4) : w4; have = true; }
  if (M_ & 2) { r = have ? cmul(r, w2) : w2; have = true; }
  if (M_ & 1) { r = have ? cmul(r, w1) : w1; have = true; }
  return r;
}
template <typename T, int R, int M_>
struct TwApply {
  __device__ static __forceinline__ void run(cx2<T>* x, const cx2<T>& w1, const cx2<T>& w2, const cx2<T>& w4,
                                             const cx2<T>& w8) {
    x[M_] = cmul(x[M_], twiddle_pow<T, M_>(w1, w2, w4, w8));
    TwApply<T, R, M_ + 1>::run(x, w1, w2, w4, w8);
  }
};
template <typename T, int R>
struct TwApply<T, R, R> {
  __device__ static __forceinline__ void run(cx2<T>*, const cx2<T>&, const cx2<T>&, const cx2<T>&, const cx2<T>&) {}
};
template <typename T, int R>
__device__ __forceinline__ void apply_twiddles(cx2<T>* x, cx2<T> w1) {
  cx2<T> w2 = w1, w4 = w1, w8 = w1;
  if (R > 2) w2 = cmul(w1, w1);
  if (R > 4) w4 = cmul(w2, w2);
  if (R > 8) w8 = cmul(w4, w4);
  TwApply<T, R, 1>::run(x, w1, w2, w4, w8);
}

template <typename T, int R, int C, bool DIT>
__device__ __forceinline__ void fft_stage(cx2<T>* __restrict__ s, const cx2<T>* __restrict__ tw, int N, int L,
                                          int tid, int nthr) {
  const int M = L / R;
  const int tstride = N / L;
  const int nbf = (N / R) * C;
  const bool pow2 = (M & (M - 1)) == 0;
  const int lgM = 31 - __clz(M);
  constexpr int PERIOD = sizeof(T) == 4 ? 16 : 8;
  const bool lin = ((M * C) % PERIOD) == 0;
  const int pstride = M * C + (M * C) / PERIOD;
  // the twiddle of the NEXT butterfly is fetched (global memory, L2 latency) while this one is computed
  auto k_of = [&](int w) {
    const int bk = (C == 1) ? w : (w / C);
    const int g = pow2 ? (bk >> lgM) : (bk / M);
    return bk - g * M;
  };
  cx2<T> tw_next = tid < nbf ? tw[k_of(tid) * tstride] : cx2<T>{(T)1, (T)0};
  for (int w = tid; w < nbf; w += nthr) {
    const cx2<T> tw_cur = tw_next;
    if (w + nthr < nbf) tw_next = tw[k_of(w + nthr) * tstride];
    const int c = (C == 1) ? 0 : (w % C);
    const int bk = (C == 1) ? w : (w / C);
    const int g = pow2 ? (bk >> lgM) : (bk / M);
    const int k = bk - g * M;
    const int e0 = (g * L + k) * C + c;
    cx2<T> x[R];
    // padded address of e0 + m*M*C: when the stride is a multiple of the padding period the pad term
    // is linear in m, so one padded base + a constant padded stride replaces R index computations
    const int a0 = fft_pad<T>(e0);
    if (lin) {
#pragma unroll
      for (int m = 0; m < R; ++m) x[m] = s[a0 + m * pstride];
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) x[m] = s[fft_pad<T>(e0 + m * M * C)];
    }
    if (DIT && k != 0) apply_twiddles<T, R>(x, tw_cur);
    SmallDft<T, R>::run(x);
    if (!DIT && k != 0) apply_twiddles<T, R>(x, tw_cur);
    if (lin) {
#pragma unroll
      for (int m = 0; m < R; ++m) s[a0 + m * pstride] = x[m];
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) s[fft_pad<T>(e0 + m * M * C)] = x[m];
    }
  }
}

// Power-of-two stage with the butterfly stride M = 2^LGM known at compile time.  Every element of
// a butterfly then sits at  pad(e0) + pad(m*M*C):  with S = R*M*C the butterfly span, e0 = A + r,
// A a multiple of S and r < M*C, the pad term (e >> 4|3) splits exactly because M*C and S are
// powers of two.  All shared-memory accesses become one base register + immediate offsets and the
// index arithmetic shrinks to shifts (the generic stage spends ~40 % of its instructions on it).
template <typename T, int R, int LGM, int C, bool DIT>
__device__ __forceinline__ void fft_stage_p2(cx2<T>* __restrict__ s, const cx2<T>* __restrict__ tw, int N, int tid,
                                             int nthr) {
  constexpr int M = 1 << LGM, MC = M * C;
  constexpr int LGC = C == 1 ? 0 : (C == 2 ? 1 : 2);
  constexpr int LGR = R == 2 ? 1 : (R == 4 ? 2 : (R == 8 ? 3 : 4));
  const int tstride = N >> (LGM + LGR);  // N / L, L = R * M
  const int nbf = (N >> LGR) * C;
  // M == 1: every twiddle is 1.  Otherwise the NEXT butterfly's twiddle is fetched while this one runs.
  cx2<T> tw_next = (M > 1 && tid < nbf) ? tw[((tid >> LGC) & (M - 1)) * tstride] : cx2<T>{(T)1, (T)0};
  for (int w = tid; w < nbf; w += nthr) {
    const cx2<T> tw_cur = tw_next;
    if (M > 1 && w + nthr < nbf) tw_next = tw[(((w + nthr) >> LGC) & (M - 1)) * tstride];
    const int c = w & (C - 1);
    const int bk = w >> LGC;
    const int g = bk >> LGM, k = bk & (M - 1);
    const int e0 = (((g << (LGM + LGR)) + k) << LGC) + c;
    cx2<T>* __restrict__ b = s + fft_pad<T>(e0);
    cx2<T> x[R];
#pragma unroll
    for (int m = 0; m < R; ++m) x[m] = b[fft_pad<T>(m * MC)];
    if (M > 1 && DIT && k != 0) apply_twiddles<T, R>(x, tw_cur);
    SmallDft<T, R>::run(x);
    if (M > 1 && !DIT && k != 0) apply_twiddles<T, R>(x, tw_cur);
#pragma unroll
    for (int m = 0; m < R; ++m) b[fft_pad<T>(m * MC)] = x[m];
  }
}

template <typename T, int R, int C, bool DIT>
__device__ __forceinline__ bool fft_stage_p2_dispatch(cx2<T>* s, const cx2<T>* tw, int N, int M, int tid, int nthr) {
  switch (M) {
    case 1: fft_stage_p2<T, R, 0, C, DIT>(s, tw, N, tid, nthr); return true;
    case 2: fft_stage_p2<T, R, 1, C, DIT>(s, tw, N, tid, nthr); return true;
    case 4: fft_stage_p2<T, R, 2, C, DIT>(s, tw, N, tid, nthr); return true;
    case 8: fft_stage_p2<T, R, 3, C, DIT>(s, tw, N, tid, nthr); return true;
    case 16: fft_stage_p2<T, R, 4, C, DIT>(s, tw, N, tid, nthr); return true;
    case 32: fft_stage_p2<T, R, 5, C, DIT>(s, tw, N, tid, nthr); return true;
    case 64: fft_stage_p2<T, R, 6, C, DIT>(s, tw, N, tid, nthr); return true;
    case 128: fft_stage_p2<T, R, 7, C, DIT>(s, tw, N, tid, nthr); return true;
    case 256: fft_stage_p2<T, R, 8, C, DIT>(s, tw, N, tid, nthr); return true;
    case 512: fft_stage_p2<T, R, 9, C, DIT>(s, tw, N, tid, nthr); return true;
    case 1024: fft_stage_p2<T, R, 10, C, DIT>(s, tw, N, tid, nthr); return true;
    default: return false;
  }
}

template <typename T, int C, bool DIT>
__device__ __forceinline__ void fft_stage_dispatch(int r, cx2<T>* s, const cx2<T>* tw, int N, int L, int tid, int nthr) {
#ifndef FFT_NO_P2
  // factorize() puts the odd radices first, so every power-of-two stage has a power-of-two stride:
  // radix 16 with M = 2^r 16^j, and one final stage (radix 2..16) with M = 1
  // (fp32: radix 16 with M = 2^r 16^j; fp64: radix 8 with M = 2^r 8^j — the host never emits radix 16 in fp64,
  // whose butterfly would need ~250 registers)
  if constexpr (sizeof(T) == 4) {
    if (r == 16) {
      if (fft_stage_p2_dispatch<T, 16, C, DIT>(s, tw, N, L >> 4, tid, nthr)) return;
    }
  } else {
    if (r == 8 && L != r) {
      if (fft_stage_p2_dispatch<T, 8, C, DIT>(s, tw, N, L >> 3, tid, nthr)) return;
    }
  }
  if (L == r) {
    if (r == 8) { fft_stage_p2<T, 8, 0, C, DIT>(s, tw, N, tid, nthr); return; }
    if (r == 4) { fft_stage_p2<T, 4, 0, C, DIT>(s, tw, N, tid, nthr); return; }
    if (r == 2) { fft_stage_p2<T, 2, 0, C, DIT>(s, tw, N, tid, nthr); return; }
  }
#endif
  if constexpr (sizeof(T) == 4) {
    if (r == 16) { fft_stage<T, 16, C, DIT>(s, tw, N, L, tid, nthr); return; }
  }
  switch (r) {
    case 8: fft_stage<T, 8, C, DIT>(s, tw, N, L, tid, nthr); break;
    case 4: fft_stage<T, 4, C, DIT>(s, tw, N, L, tid, nthr); break;
    case 2: fft_stage<T, 2, C, DIT>(s, tw, N, L, tid, nthr); break;
    case 3: fft_stage<T, 3, C, DIT>(s, tw, N, L, tid, nthr); break;
    case 5: fft_stage<T, 5, C, DIT>(s, tw, N, L, tid, nthr); break;
    case 7: fft_stage<T, 7, C, DIT>(s, tw, N, L, tid, nthr); break;
    default: fft_stage<T, 11, C, DIT>(s, tw, N, L, tid, nthr); break;
  }
}

// Forward FFT of the C interleaved length-N arrays in shared memory.  The caller must
// __syncthreads() after filling `s`; on return all stages are complete and synchronised.
template <typename T, int C>
__device__ __forceinline__ void fft_dif(cx2<T>* s, const cx2<T>* tw, const FftDesc& d, int tid, int nthr) {
  int L = d.n;
  for (int st = 0; st < d.nstage; ++st) {
    fft_stage_dispatch<T, C, false>(d.radix[st], s, tw, d.n, L, tid, nthr);
    L /= d.radix[st];
    __syncthreads();
  }
}
template <typename T, int C>
__device__ __forceinline__ void fft_dit(cx2<T>* s, const cx2<T>* tw, const FftDesc& d, int tid, int nthr) {
  int L = 1;
  for (int st = d.nstage - 1; st >= 0; --st) {
    L *= d.radix[st];
    fft_stage_dispatch<T, C, true>(d.radix[st], s, tw, d.n, L, tid, nthr);
    __syncthreads();
  }
}

// debug / unit-test kernel: one CTA per transform, natural order in and out.
//   mode 0: DIF (output read back through rev[]),  mode 1: DIT (input scattered through posmap[]).
template <typename T>
__global__ void k_fft_debug(FftDesc d, const cx2<T>* __restrict__ tw, const int* __restrict__ rev,
                            const int* __restrict__ posmap, const cx2<T>* __restrict__ in, cx2<T>* __restrict__ out,
                            int mode, int inverse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx2<T>* s = reinterpret_cast<cx2<T>*>(smem_raw);
  const int N = d.n, tid = threadIdx.x, nthr = blockDim.x;
  const cx2<T>* src = in + (size_t)blockIdx.x * N;
  cx2<T>* dst = out + (size_t)blockIdx.x * N;
  const T sg = inverse ? (T)-1 : (T)1;
  for (int n = tid; n < N; n += nthr) {
    cx2<T> v = src[n];
    v.y *= sg;
    s[fft_pad<T>(mode ? posmap[n] : n)] = v;
  }
  __syncthreads();
  if (mode) fft_dit<T, 1>(s, tw, d, tid, nthr);
  else fft_dif<T, 1>(s, tw, d, tid, nthr);
  for (int p = tid; p < N; p += nthr) {
    cx2<T> v = s[fft_pad<T>(p)];
    v.y *= sg;
    dst[mode ? p : rev[p]] = v;
  }
}
