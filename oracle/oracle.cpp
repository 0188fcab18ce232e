// oracle.cpp — CPU restatement of the go-tfhe gate-bootstrap hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
// (go-tfhe_b200/) never links, imports or calls anything in oracle/.
//
// Parity status: the reference (thedonutfactory/go-tfhe @ 9c41b7a) is pure Go and there is
// no Go toolchain in this image, so the reference itself cannot be run (no oracle/_ref).
// The reference holds NO ciphertext-level golden vectors (every RNG is unseeded
// math/rand: key/key.go:17, tlwe/tlwe.go:37), so at ciphertext level this oracle is
// "PARITY UNPINNED".  What the reference's tests do pin, and what tests/test_oracle_*.py check
// this file against, are plaintext-level facts: torus constants (utils/utils_test.go:15-19),
// FFT round trip <= 10 LSB (poly/poly_test.go:24-32), every gate truth table
// (gates/gates_test.go:27-353), batch AND/OR/XOR (gates_test.go:369-480), PBS
// identity/NOT/constant/LUT-reuse (evaluator/programmable_bootstrap_test.go:30-187) and
// Uint identity/complement/modulo (params/uint_params_test.go:61-126).  Independently of the
// reference, tests also check the polynomial product against an exact big-integer negacyclic
// convolution.
//
// Arithmetic fidelity: compile with -O2 -ffp-contract=off (Go/amd64 does not fuse FMA);
// math.Round == std::round (half away from zero); Torus(int64(x)) == truncating cast then
// wrap to 32 bits.  Twiddles use libm sincos where Go uses math.Sincos (<= 1 ulp apart).
// Randomness: the reference has no seed semantics; this file uses xoshiro256** + Box-Muller.
//
// Each function cites the reference file:line it follows (paths relative to /root/reference).

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

typedef uint32_t Torus;  // params/params.go:27

extern "C" {
struct oracle_params {
  int32_t n;        // TLWELv0.N
  int32_t N;        // TRGSWLv1.N
  int32_t nbit;     // TRGSWLv1.NBIT
  int32_t bgbit;    // TRGSWLv1.BGBIT
  int32_t L;        // TRGSWLv1.L
  int32_t basebit;  // TRGSWLv1.BASEBIT
  int32_t iks_t;    // TRGSWLv1.IKS_T
  int32_t _pad;
  double alpha_lv0; // TLWELv0.ALPHA  (== KSKAlpha, params.go:629)
  double alpha_lv1; // TLWELv1.ALPHA  (== BSKAlpha, params.go:634)
};
}

// ---------------------------------------------------------------------------------------------
// Parameter sets — params/params.go:83-391 (Uint6-8 omitted: the reference marks them broken,
// params/UINT_STATUS.md:14-31).
// ---------------------------------------------------------------------------------------------
struct NamedParams { const char* name; oracle_params p; };
static const NamedParams kParamSets[] = {
    {"80",    {550, 1024, 10, 6, 3, 2, 7, 0, 5.0e-5, 3.73e-8}},                                   // :83-112
    {"110",   {630, 1024, 10, 6, 3, 2, 8, 0, 3.0517578125e-05, 2.980232238769531e-8}},            // :117-146
    {"128",   {700, 1024, 10, 6, 3, 2, 9, 0, 2.0e-5, 2.0e-8}},                                    // :151-180
    {"uint1", {700, 1024, 10, 10, 2, 2, 8, 0, 2.0e-05, 2.0e-08}},                                 // :194-223
    {"uint2", {687, 512, 9, 18, 1, 4, 3, 0, 0.00002120846893069971872305794214,
               0.00000000000231841227527049948463}},                                              // :236-265
    {"uint3", {820, 1024, 10, 23, 1, 6, 2, 0, 0.00000251676160959795544987084234,
               0.00000000000000022204460492503131}},                                              // :277-306
    {"uint4", {820, 2048, 11, 22, 1, 5, 3, 0, 0.00000251676160959795544987084234,
               0.00000000000000022204460492503131}},                                              // :318-347
    {"uint5", {1071, 2048, 11, 22, 1, 6, 3, 0, 7.088226765410429399593757e-08,
               2.2204460492503131e-17}},                                                          // :362-391
};

// ---------------------------------------------------------------------------------------------
// RNG (no reference equivalent: math/rand is unseeded there)
// ---------------------------------------------------------------------------------------------
struct Rng {
  uint64_t s[4];
  bool have_spare = false;
  double spare = 0.0;
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  uint32_t u32() { return (uint32_t)(next() >> 32); }
  double uniform() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double normal() {  // stands in for rng.NormFloat64()
    if (have_spare) { have_spare = false; return spare; }
    double u = uniform(), v = uniform();
    double r = std::sqrt(-2.0 * std::log(u));
    spare = r * std::sin(2.0 * M_PI * v); have_spare = true;
    return r * std::cos(2.0 * M_PI * v);
  }
};

// ---------------------------------------------------------------------------------------------
// utils/utils.go
// ---------------------------------------------------------------------------------------------
static Torus f64_to_torus(double d) {  // utils/utils.go:11-14
  double t = std::fmod(d, 1.0) * 4294967296.0;
  return (Torus)(uint64_t)(int64_t)t;
}
static Torus gaussian_f64(double mu, double stddev, Rng& rng) {  // utils/utils.go:31-41
  Torus mu_t = f64_to_torus(mu);
  double sample = rng.normal() * stddev;
  return mu_t + f64_to_torus(sample);
}

// ---------------------------------------------------------------------------------------------
// poly.Evaluator — twiddles (poly/poly_evaluator.go:76-165) and transforms
// (poly/fourier_transform.go).  A FourierPoly is N doubles laid out in groups of eight:
// four real parts then four imaginary parts (poly/poly.go:54-62).
// ---------------------------------------------------------------------------------------------
typedef std::complex<double> cplx;

static cplx go_cexp_i(double e) {  // cmplx.Exp(complex(0, e)) == (cos e, sin e)
  return cplx(std::cos(e), std::sin(e));
}
static cplx go_cmul(cplx a, cplx b) {  // Go complex128 product, no FMA
  return cplx(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
template <class T> static void bit_reverse_in_place(std::vector<T>& d) {  // poly_evaluator.go:146-165
  size_t n = d.size();
  if (n <= 1) return;
  size_t j = 0;
  for (size_t i = 0; i < n; i++) {
    if (i < j) std::swap(d[i], d[j]);
    size_t m = n >> 1;
    while (m > 0 && j >= m) { j -= m; m >>= 1; }
    j += m;
  }
}

struct PolyEval {
  int N;                       // polynomial degree
  std::vector<cplx> tw, twInv; // poly_evaluator.go:114-143 (length N/2 - 1 each)
  std::vector<double> fa, fb, facc, fbcc;
  std::vector<std::vector<double>> decFFT;
  std::vector<std::vector<Torus>> dec;
  std::vector<Torus> tmpA, tmpB, resA, resB, acc1A, acc1B, acc2A, acc2B;

  explicit PolyEval(int N_) : N(N_) {
    gen_twiddles(N / 2);
    fa.resize(N); fb.resize(N); facc.resize(N); fbcc.resize(N);
    decFFT.assign(8, std::vector<double>(N)); dec.assign(8, std::vector<Torus>(N));
    for (auto* v : {&tmpA, &tmpB, &resA, &resB, &acc1A, &acc1B, &acc2A, &acc2B}) v->resize(N);
  }

  void gen_twiddles(int M) {  // genTwiddleFactors(N/2), poly_evaluator.go:114-143
    std::vector<cplx> twFFT(M / 2), twInvFFT(M / 2);
    for (int i = 0; i < M / 2; i++) {
      double e = -2 * M_PI * (double)i / (double)M;
      twFFT[i] = go_cexp_i(e);
      twInvFFT[i] = go_cexp_i(-e);
    }
    bit_reverse_in_place(twFFT);
    bit_reverse_in_place(twInvFFT);
    for (int m = 1, t = M / 2; m <= M / 2; m <<= 1, t >>= 1) {
      cplx fold = go_cexp_i(2 * M_PI * (double)t / (double)(4 * M));
      for (int i = 0; i < m; i++) tw.push_back(go_cmul(twFFT[i], fold));
    }
    for (int m = M / 2, t = 1; m >= 1; m >>= 1, t <<= 1) {
      cplx fold = go_cexp_i(-2 * M_PI * (double)t / (double)(4 * M));
      for (int i = 0; i < m; i++) twInv.push_back(go_cmul(twInvFFT[i], fold));
    }
  }

  // fourier_transform.go:64-85 — fold p[j] + i*p[j+N/2] into the 4re/4im layout.
  void fold_in(const Torus* p, double* fp) const {
    for (int i = 0, ii = 0; i < N; i += 8, ii += 4)
      for (int k = 0; k < 4; k++) {
        fp[i + k] = (double)(int32_t)p[ii + k];
        fp[i + 4 + k] = (double)(int32_t)p[ii + N / 2 + k];
      }
  }
  // fourier_transform.go:170-174
  static inline void bfly(double& uR, double& uI, double& vR, double& vI, double wR, double wI) {
    double vwR = vR * wR - vI * wI;
    double vwI = vR * wI + vI * wR;
    double a = uR + vwR, b = uI + vwI, c = uR - vwR, d = uI - vwI;
    uR = a; uI = b; vR = c; vI = d;
  }
  // fourier_transform.go:250-255
  static inline void ibfly(double& uR, double& uI, double& vR, double& vI, double wR, double wI) {
    double sR = uR + vR, sI = uI + vI, dR = uR - vR, dI = uI - vI;
    uR = sR; uI = sI;
    vR = dR * wR - dI * wI;
    vI = dR * wI + dI * wR;
  }
  // fourier_transform.go:178-247.  c has N doubles (N/2 complex values in 4re/4im groups).
  void fft_in_place(double* c) const {
    int w = 0;
    {  // first stage: one twiddle, partner half an array away
      double wR = tw[w].real(), wI = tw[w].imag(); w++;
      for (int j = 0; j < N / 2; j += 8)
        for (int k = 0; k < 4; k++) bfly(c[j + k], c[j + 4 + k], c[j + N / 2 + k], c[j + N / 2 + 4 + k], wR, wI);
    }
    int t = N / 2;
    for (int m = 2; m <= N / 16; m <<= 1) {  // middle stages, whole groups of eight
      t >>= 1;
      for (int i = 0; i < m; i++) {
        int j1 = 2 * i * t, j2 = j1 + t;
        double wR = tw[w].real(), wI = tw[w].imag(); w++;
        for (int j = j1; j < j2; j += 8)
          for (int k = 0; k < 4; k++) bfly(c[j + k], c[j + 4 + k], c[j + t + k], c[j + t + 4 + k], wR, wI);
      }
    }
    for (int j = 0; j < N; j += 8) {  // second-to-last: lanes (0,2) (1,3) inside a group
      double wR = tw[w].real(), wI = tw[w].imag(); w++;
      double* re = c + j; double* im = c + j + 4;
      bfly(re[0], im[0], re[2], im[2], wR, wI);
      bfly(re[1], im[1], re[3], im[3], wR, wI);
    }
    for (int j = 0; j < N; j += 8) {  // last: lanes (0,1) (2,3), two twiddles per group
      double wR0 = tw[w].real(), wI0 = tw[w].imag(), wR1 = tw[w + 1].real(), wI1 = tw[w + 1].imag(); w += 2;
      double* re = c + j; double* im = c + j + 4;
      bfly(re[0], im[0], re[1], im[1], wR0, wI0);
      bfly(re[2], im[2], re[3], im[3], wR1, wI1);
    }
  }
  // fourier_transform.go:258-347
  void ifft_in_place(double* c) const {
    int w = 0;
    for (int j = 0; j < N; j += 8) {
      double wR0 = twInv[w].real(), wI0 = twInv[w].imag(), wR1 = twInv[w + 1].real(), wI1 = twInv[w + 1].imag(); w += 2;
      double* re = c + j; double* im = c + j + 4;
      ibfly(re[0], im[0], re[1], im[1], wR0, wI0);
      ibfly(re[2], im[2], re[3], im[3], wR1, wI1);
    }
    for (int j = 0; j < N; j += 8) {
      double wR = twInv[w].real(), wI = twInv[w].imag(); w++;
      double* re = c + j; double* im = c + j + 4;
      ibfly(re[0], im[0], re[2], im[2], wR, wI);
      ibfly(re[1], im[1], re[3], im[3], wR, wI);
    }
    int t = 8;
    for (int m = N / 16; m >= 2; m >>= 1) {
      for (int i = 0; i < m; i++) {
        int j1 = 2 * i * t, j2 = j1 + t;
        double wR = twInv[w].real(), wI = twInv[w].imag(); w++;
        for (int j = j1; j < j2; j += 8)
          for (int k = 0; k < 4; k++) ibfly(c[j + k], c[j + 4 + k], c[j + t + k], c[j + t + 4 + k], wR, wI);
      }
      t <<= 1;
    }
    double scale = (double)(N / 2);
    double wR = twInv[w].real(), wI = twInv[w].imag();
    for (int j = 0; j < N / 2; j += 8) {
      for (int k = 0; k < 4; k++) ibfly(c[j + k], c[j + 4 + k], c[j + N / 2 + k], c[j + N / 2 + 4 + k], wR, wI);
      for (int k = 0; k < 8; k++) { c[j + k] /= scale; c[j + N / 2 + k] /= scale; }
    }
  }
  // fourier_transform.go:88-104
  void float_mod_q(double* c) const {
    const double Q = 4294967296.0;
    for (int i = 0; i < N; i++) c[i] = std::round(c[i] - Q * std::round(c[i] / Q));
  }
  // fourier_transform.go:107-125
  void unfold_out(const double* fp, Torus* p) const {
    for (int i = 0, ii = 0; i < N; i += 8, ii += 4)
      for (int k = 0; k < 4; k++) {
        p[ii + k] = (Torus)(uint64_t)(int64_t)fp[i + k];
        p[ii + N / 2 + k] = (Torus)(uint64_t)(int64_t)fp[i + 4 + k];
      }
  }
  void to_fourier(const Torus* p, double* fp) const { fold_in(p, fp); fft_in_place(fp); }   // :18-21
  void to_poly_unsafe(double* fp, Torus* p) const { ifft_in_place(fp); float_mod_q(fp); unfold_out(fp, p); }  // :40-44

  // poly/fourier_ops.go:138-161
  void mul_cmplx(const double* v0, const double* v1, double* out) const {
    for (int i = 0; i < N; i += 8)
      for (int k = 0; k < 4; k++) {
        double r = v0[i + k] * v1[i + k] - v0[i + 4 + k] * v1[i + 4 + k];
        double im = v0[i + k] * v1[i + 4 + k] + v0[i + 4 + k] * v1[i + k];
        out[i + k] = r; out[i + 4 + k] = im;
      }
  }
  // poly/fourier_ops.go:167-191
  void mul_add_cmplx(const double* v0, const double* v1, double* out) const {
    for (int i = 0; i < N; i += 8)
      for (int k = 0; k < 4; k++) {
        double r = out[i + k] + (v0[i + k] * v1[i + k] - v0[i + 4 + k] * v1[i + 4 + k]);
        double im = out[i + 4 + k] + (v0[i + k] * v1[i + 4 + k] + v0[i + 4 + k] * v1[i + k]);
        out[i + k] = r; out[i + 4 + k] = im;
      }
  }
  // poly/poly_mul.go:12-22
  void mul_poly(const Torus* p0, const Torus* p1, Torus* out) {
    to_fourier(p0, fa.data());
    to_fourier(p1, fb.data());
    mul_cmplx(fa.data(), fb.data(), fa.data());
    to_poly_unsafe(fa.data(), out);
  }
};

// poly/buffer_methods.go:133-164 — multiply by X^k in Z[X]/(X^N+1); "negation" is 0xFFFFFFFF - a.
static void poly_mul_xk(const Torus* a, int N, long k, Torus* out) {
  k = k % (2L * N);
  if (k == 0) { std::memcpy(out, a, sizeof(Torus) * N); return; }
  if (k < 0) k += 2L * N;
  if (k < N) {
    for (long i = 0; i < N - k; i++) out[i + k] = a[i];
    for (long i = N - k; i < N; i++) out[i + k - N] = ~(Torus)0 - a[i];
  } else {
    k -= N;
    for (long i = 0; i < N - k; i++) out[i + k] = ~(Torus)0 - a[i];
    for (long i = N - k; i < N; i++) out[i + k - N] = a[i];
  }
}

// poly/decomposer.go:55-66
static void decompose_poly(const Torus* p, int N, int bgbit, int level, Torus offset, std::vector<Torus>* out) {
  Torus mask = (Torus)((1u << bgbit) - 1), half = (Torus)(1u << (bgbit - 1));
  for (int j = 0; j < N; j++) {
    Torus tmp = p[j] + offset;
    for (int i = 0; i < level; i++) out[i][j] = ((tmp >> (32 - (i + 1) * bgbit)) & mask) - half;
  }
}

// cloudkey/cloudkey.go:60-71
static Torus decomposition_offset(const oracle_params& P) {
  Torus off = 0, bg = 1u << P.bgbit;
  for (int i = 0; i < P.L; i++) off += (Torus)(bg / 2) * (Torus)(1u << (32 - (i + 1) * P.bgbit));
  return off;
}

// evaluator/evaluator.go:50-81.  bsk_row: [2L][2][N] doubles (row r: A spectrum then B spectrum).
static void external_product(PolyEval& ev, const oracle_params& P, const double* bsk_row, const Torus* inA,
                             const Torus* inB, Torus offset, Torus* outA, Torus* outB) {
  const int N = P.N, L = P.L;
  decompose_poly(inA, N, P.bgbit, L, offset, ev.dec.data());
  decompose_poly(inB, N, P.bgbit, L, offset, ev.dec.data() + L);
  for (int i = 0; i < 2 * L; i++) ev.to_fourier(ev.dec[i].data(), ev.decFFT[i].data());
  std::fill(ev.facc.begin(), ev.facc.end(), 0.0);
  std::fill(ev.fbcc.begin(), ev.fbcc.end(), 0.0);
  for (int i = 0; i < 2 * L; i++) {
    ev.mul_add_cmplx(ev.decFFT[i].data(), bsk_row + (size_t)(2 * i) * N, ev.facc.data());
    ev.mul_add_cmplx(ev.decFFT[i].data(), bsk_row + (size_t)(2 * i + 1) * N, ev.fbcc.data());
  }
  ev.to_poly_unsafe(ev.facc.data(), outA);
  ev.to_poly_unsafe(ev.fbcc.data(), outB);
}

// evaluator/evaluator.go:85-106 — out = ct0 + cond (x) (ct1 - ct0); out may alias ct0.
static void cmux(PolyEval& ev, const oracle_params& P, const double* bsk_row, const Torus* c0A, const Torus* c0B,
                 const Torus* c1A, const Torus* c1B, Torus offset, Torus* outA, Torus* outB) {
  const int N = P.N;
  for (int i = 0; i < N; i++) { ev.tmpA[i] = c1A[i] - c0A[i]; ev.tmpB[i] = c1B[i] - c0B[i]; }
  if (outA != c0A) { std::memcpy(outA, c0A, 4 * N); std::memcpy(outB, c0B, 4 * N); }
  external_product(ev, P, bsk_row, ev.tmpA.data(), ev.tmpB.data(), offset, ev.resA.data(), ev.resB.data());
  for (int i = 0; i < N; i++) { outA[i] += ev.resA[i]; outB[i] += ev.resB[i]; }
}

// evaluator/evaluator.go:110-135.  ct: n+1 words; testvec: [2][N]; bsk: [n][2L][2][N]; out: [2][N].
static void blind_rotate(PolyEval& ev, const oracle_params& P, const Torus* ct, const Torus* testvec,
                         const double* bsk, Torus offset, Torus* out) {
  const int N = P.N, nbit = P.nbit, n = P.n;
  long b_tilda = 2L * N - (((long)ct[n] + (1L << (31 - nbit - 1))) >> (32 - nbit - 1));
  poly_mul_xk(testvec, N, b_tilda, ev.acc1A.data());
  poly_mul_xk(testvec + N, N, b_tilda, ev.acc1B.data());
  const size_t row = (size_t)2 * P.L * 2 * N;
  for (int i = 0; i < n; i++) {
    long a_tilda = (long)((Torus)(ct[i] + (Torus)(1u << (31 - nbit - 1))) >> (32 - nbit - 1));
    poly_mul_xk(ev.acc1A.data(), N, a_tilda, ev.acc2A.data());
    poly_mul_xk(ev.acc1B.data(), N, a_tilda, ev.acc2B.data());
    cmux(ev, P, bsk + row * i, ev.acc1A.data(), ev.acc1B.data(), ev.acc2A.data(), ev.acc2B.data(), offset,
         ev.acc1A.data(), ev.acc1B.data());
  }
  std::memcpy(out, ev.acc1A.data(), 4 * N);
  std::memcpy(out + N, ev.acc1B.data(), 4 * N);
}

// trlwe/trlwe_ops.go:10-21 with k = 0.
static void sample_extract0(const Torus* trlwe, int N, Torus* out) {
  const int k = 0;
  for (int i = 0; i < N; i++) out[i] = (i <= k) ? trlwe[k - i] : ~(Torus)0 - trlwe[N + k - i];
  out[N] = trlwe[N + k];
}

// trgsw/keyswitch.go:10-37.  ksk: [N][t][base][n+1], row idx = base*t*i + base*j + k.
static void key_switch(const oracle_params& P, const Torus* src, const Torus* ksk, Torus* out) {
  const int N = P.N, basebit = P.basebit, base = 1 << basebit, t = P.iks_t, n = P.n;
  for (int x = 0; x <= n; x++) out[x] = 0;
  out[n] = src[N];
  Torus prec = (Torus)(1u << (32 - (1 + basebit * t)));
  for (int i = 0; i < N; i++) {
    Torus abar = src[i] + prec;
    for (int j = 0; j < t; j++) {
      Torus k = (abar >> (32 - (j + 1) * basebit)) & (Torus)((1u << basebit) - 1);
      if (k != 0) {
        const Torus* rowp = ksk + ((size_t)base * t * i + (size_t)base * j + k) * (n + 1);
        for (int x = 0; x <= n; x++) out[x] -= rowp[x];
      }
    }
  }
}

// evaluator/evaluator.go:139-148 (and programmable_bootstrap.go:93-115 when testvec is a LUT).
static void bootstrap(PolyEval& ev, const oracle_params& P, const Torus* ct, const Torus* testvec, const double* bsk,
                      const Torus* ksk, Torus offset, Torus* out) {
  std::vector<Torus> rot(2 * P.N), ext(P.N + 1);
  blind_rotate(ev, P, ct, testvec, bsk, offset, rot.data());
  sample_extract0(rot.data(), P.N, ext.data());
  key_switch(P, ext.data(), ksk, out);
}

// Gate prologues.  op codes are shared with include/tfhe_b200.h.
// evaluator/gates_helper.go:10-63 (NAND/AND/OR/XOR), gates/gates.go:52-104 (the rest).
enum { OP_NAND = 0, OP_AND, OP_OR, OP_XOR, OP_XNOR, OP_NOR, OP_ANDNY, OP_ANDYN, OP_ORNY, OP_ORYN };
static int gate_prepare(const oracle_params& P, int op, const Torus* a, const Torus* b, Torus* out) {
  const int n = P.n;
  const Torus e8 = f64_to_torus(0.125), m8 = f64_to_torus(-0.125), q4 = f64_to_torus(0.25);
  Torus bias;
  switch (op) {
    case OP_NAND:  for (int i = 0; i <= n; i++) out[i] = (Torus)0 - (a[i] + b[i]); bias = e8; break;   // helper:10-21
    case OP_AND:   for (int i = 0; i <= n; i++) out[i] = a[i] + b[i]; bias = m8; break;                // helper:24-35
    case OP_OR:    for (int i = 0; i <= n; i++) out[i] = a[i] + b[i]; bias = e8; break;                // helper:38-49
    case OP_XOR:   for (int i = 0; i <= n; i++) out[i] = a[i] + 2 * b[i]; bias = q4; break;            // helper:52-63
    case OP_XNOR:  for (int i = 0; i <= n; i++) out[i] = a[i] - b[i] * 2; bias = q4; break;            // gates.go:52-58
    case OP_NOR:   for (int i = 0; i <= n; i++) out[i] = (Torus)0 - (a[i] + b[i]); bias = m8; break;   // gates.go:72-76
    case OP_ANDNY: for (int i = 0; i <= n; i++) out[i] = ((Torus)0 - a[i]) + b[i]; bias = m8; break;   // gates.go:79-83
    case OP_ANDYN: for (int i = 0; i <= n; i++) out[i] = a[i] - b[i]; bias = m8; break;                // gates.go:86-90
    case OP_ORNY:  for (int i = 0; i <= n; i++) out[i] = ((Torus)0 - a[i]) + b[i]; bias = e8; break;   // gates.go:93-97
    case OP_ORYN:  for (int i = 0; i <= n; i++) out[i] = a[i] - b[i]; bias = e8; break;                // gates.go:100-104
    default: return -1;
  }
  out[n] += bias;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Client side (needed only to make test data)
// ---------------------------------------------------------------------------------------------
// tlwe/tlwe.go:36-51
static void lwe_encrypt_f64(const oracle_params& P, double p, double alpha, const Torus* key, Rng& rng, Torus* out) {
  Torus inner = 0;
  for (int i = 0; i < P.n; i++) { Torus r = rng.u32(); inner += key[i] * r; out[i] = r; }
  out[P.n] = inner + gaussian_f64(p, alpha, rng);
}
// trlwe/trlwe.go:28-50
static void trlwe_encrypt_f64(PolyEval& ev, const oracle_params& P, const double* p, double alpha, const Torus* key,
                              Rng& rng, Torus* A, Torus* B) {
  const int N = P.N;
  for (int i = 0; i < N; i++) A[i] = rng.u32();
  for (int i = 0; i < N; i++) B[i] = gaussian_f64(p ? p[i] : 0.0, alpha, rng);
  std::vector<Torus> prod(N);
  ev.mul_poly(A, key, prod.data());
  for (int i = 0; i < N; i++) B[i] += prod[i];
}
// trgsw/trgsw.go:32-57 then :71-82 — one TRGSW, returned directly in Fourier form [2L][2][N].
static void trgsw_encrypt_fft(PolyEval& ev, const oracle_params& P, Torus p, double alpha, const Torus* key, Rng& rng,
                              double* out) {
  const int N = P.N, L = P.L;
  std::vector<std::vector<Torus>> A(2 * L, std::vector<Torus>(N)), B(2 * L, std::vector<Torus>(N));
  for (int i = 0; i < 2 * L; i++) trlwe_encrypt_f64(ev, P, nullptr, alpha, key, rng, A[i].data(), B[i].data());
  double bg = (double)(1u << P.bgbit);
  for (int i = 0; i < L; i++) {
    Torus g = f64_to_torus(1.0 / std::pow(bg, (double)(i + 1)));
    A[i][0] += p * g;
    B[i + L][0] += p * g;
  }
  for (int i = 0; i < 2 * L; i++) {
    ev.to_fourier(A[i].data(), out + (size_t)(2 * i) * N);
    ev.to_fourier(B[i].data(), out + (size_t)(2 * i + 1) * N);
  }
}

template <class F> static void parallel_for(int count, int threads, F f) {
  if (threads < 1) threads = 1;
  if (threads > count) threads = count > 0 ? count : 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([=]() { for (int i = t; i < count; i += threads) f(i, t); });
  for (auto& th : pool) th.join();
}

// ---------------------------------------------------------------------------------------------
// C interface (ctypes)
// ---------------------------------------------------------------------------------------------
extern "C" {

int oracle_get_params(const char* name, oracle_params* out) {
  for (auto& s : kParamSets)
    if (std::strcmp(s.name, name) == 0) { *out = s.p; return 0; }
  return -1;
}
uint32_t oracle_f64_to_torus(double d) { return f64_to_torus(d); }
uint32_t oracle_decomposition_offset(const oracle_params* P) { return decomposition_offset(*P); }

void* oracle_eval_new(int N) { return new PolyEval(N); }
void oracle_eval_free(void* e) { delete (PolyEval*)e; }
int oracle_twiddle_count(void* e) { return (int)((PolyEval*)e)->tw.size(); }
void oracle_get_twiddles(void* e, double* tw, double* twInv) {
  auto* ev = (PolyEval*)e;
  for (size_t i = 0; i < ev->tw.size(); i++) {
    tw[2 * i] = ev->tw[i].real(); tw[2 * i + 1] = ev->tw[i].imag();
    twInv[2 * i] = ev->twInv[i].real(); twInv[2 * i + 1] = ev->twInv[i].imag();
  }
}
void oracle_to_fourier(void* e, const uint32_t* p, double* fp) { ((PolyEval*)e)->to_fourier(p, fp); }
void oracle_to_poly(void* e, const double* fp, uint32_t* p) {  // ToPolyAssign, fourier_transform.go:31-36
  auto* ev = (PolyEval*)e;
  std::vector<double> tmp(fp, fp + ev->N);
  ev->to_poly_unsafe(tmp.data(), p);
}
void oracle_mul_poly(void* e, const uint32_t* a, const uint32_t* b, uint32_t* out) { ((PolyEval*)e)->mul_poly(a, b, out); }
void oracle_poly_mul_xk(const uint32_t* a, int N, int64_t k, uint32_t* out) { poly_mul_xk(a, N, (long)k, out); }
void oracle_decompose(const oracle_params* P, const uint32_t* p, uint32_t offset, uint32_t* out /*[L][N]*/) {
  std::vector<std::vector<Torus>> d(P->L, std::vector<Torus>(P->N));
  decompose_poly(p, P->N, P->bgbit, P->L, offset, d.data());
  for (int i = 0; i < P->L; i++) std::memcpy(out + (size_t)i * P->N, d[i].data(), 4 * P->N);
}

// key/key.go:16-45 (binary keys)
void oracle_secret_key(const oracle_params* P, uint64_t seed, uint32_t* s0, uint32_t* s1) {
  Rng rng(seed);
  for (int i = 0; i < P->n; i++) s0[i] = (rng.next() >> 63) ? 1 : 0;
  for (int i = 0; i < P->N; i++) s1[i] = (rng.next() >> 63) ? 1 : 0;
}

// tlwe/tlwe.go:54-62 (EncryptBool), :36-51 (EncryptF64); count ciphertexts from one seeded stream.
void oracle_encrypt_bool(const oracle_params* P, const uint32_t* s0, uint64_t seed, int count, const uint8_t* bits,
                         uint32_t* out) {
  Rng rng(seed);
  for (int g = 0; g < count; g++)
    lwe_encrypt_f64(*P, bits[g] ? 0.125 : -0.125, P->alpha_lv0, s0, rng, out + (size_t)g * (P->n + 1));
}
// tlwe/tlwe.go:65-74
void oracle_decrypt_bool(const oracle_params* P, const uint32_t* s0, int count, const uint32_t* ct, uint8_t* bits) {
  for (int g = 0; g < count; g++) {
    const Torus* c = ct + (size_t)g * (P->n + 1);
    Torus inner = 0;
    for (int i = 0; i < P->n; i++) inner += c[i] * s0[i];
    bits[g] = ((int32_t)(c[P->n] - inner) >= 0) ? 1 : 0;
  }
}
// raw phase b - <a,s>, for noise / tolerance checks (same inner product as tlwe.go:66-71)
void oracle_phase(const oracle_params* P, const uint32_t* s0, int count, const uint32_t* ct, uint32_t* phase) {
  for (int g = 0; g < count; g++) {
    const Torus* c = ct + (size_t)g * (P->n + 1);
    Torus inner = 0;
    for (int i = 0; i < P->n; i++) inner += c[i] * s0[i];
    phase[g] = c[P->n] - inner;
  }
}
// tlwe/programmable_encrypt.go:12-27
void oracle_encrypt_message(const oracle_params* P, const uint32_t* s0, uint64_t seed, int count, const int32_t* msgs,
                            int msg_mod, uint32_t* out) {
  Rng rng(seed);
  double scale = (double)(1ull << 31) / (double)msg_mod;
  for (int g = 0; g < count; g++) {
    int m = msgs[g] % msg_mod;
    if (m < 0) m += msg_mod;
    double enc = (double)m * scale / (double)(1ull << 32);
    lwe_encrypt_f64(*P, enc, P->alpha_lv0, s0, rng, out + (size_t)g * (P->n + 1));
  }
}
// tlwe/programmable_encrypt.go:33-54
void oracle_decrypt_message(const oracle_params* P, const uint32_t* s0, int count, const uint32_t* ct, int msg_mod,
                            int32_t* msgs) {
  Torus scale = (Torus)(1ull << 31) / (Torus)msg_mod;
  for (int g = 0; g < count; g++) {
    const Torus* c = ct + (size_t)g * (P->n + 1);
    Torus inner = 0;
    for (int i = 0; i < P->n; i++) inner += c[i] * s0[i];
    Torus phase = c[P->n] - inner;
    int decoded = (int)((Torus)(phase + scale / 2) / scale);
    int m = decoded % msg_mod;
    if (m < 0) m += msg_mod;
    msgs[g] = m;
  }
}

// lut/generator.go:56-100 with lut/encoder.go:47-58; fvals[x] = f(x) for x in [0, msg_mod).
void oracle_gen_lut(const oracle_params* P, int msg_mod, const int32_t* fvals, uint32_t* out /*[2][N]*/) {
  const int N = P->N;
  auto div_round = [](long a, long b) { return (a + b / 2) / b; };  // generator.go:170-173
  std::vector<Torus> raw(N, 0), rot(N);
  double enc_scale = 1.0 / (double)(2 * msg_mod);  // encoder.go:21-27
  for (int x = 0; x < msg_mod; x++) {
    long start = div_round((long)x * N, msg_mod), end = div_round((long)(x + 1) * N, msg_mod);
    int y = fvals[x] % msg_mod;
    if (y < 0) y += msg_mod;
    Torus ey = f64_to_torus((double)y * enc_scale);
    for (long xx = start; xx < end; xx++) raw[xx] = ey;
  }
  long offset = div_round(N, 2L * msg_mod);
  for (int i = 0; i < N; i++) rot[i] = raw[(i + offset) % N];
  for (long i = N - offset; i < N; i++) rot[i] = (Torus)0 - rot[i];
  for (int i = 0; i < N; i++) { out[i] = 0; out[N + i] = rot[i]; }
}

// cloudkey/cloudkey.go:24-31,60-145.  Outputs: testvec [2][N]; ksk [N][t][base][n+1] (k = 0 rows stay
// zero, :104-106); bsk_fft [n][2L][2][N] doubles in the reference FourierPoly layout.
void oracle_cloudkey(const oracle_params* Pp, const uint32_t* s0, const uint32_t* s1, uint64_t seed, int threads,
                     uint32_t* offset_out, uint32_t* testvec, uint32_t* ksk, double* bsk_fft) {
  const oracle_params P = *Pp;
  *offset_out = decomposition_offset(P);
  Torus e8 = f64_to_torus(0.125);
  for (int i = 0; i < P.N; i++) { testvec[i] = 0; testvec[P.N + i] = e8; }   // :74-85
  const int base = 1 << P.basebit, t = P.iks_t;
  if (ksk) {
    parallel_for(P.N, threads, [&](int i, int) {                               // :88-120
      Rng rng(seed * 0x100000001B3ull + 0x1000000ull + (uint64_t)i);
      for (int j = 0; j < t; j++)
        for (int k = 0; k < base; k++) {
          Torus* row = ksk + ((size_t)base * t * i + (size_t)base * j + k) * (P.n + 1);
          if (k == 0) { std::memset(row, 0, 4 * (P.n + 1)); continue; }
          unsigned shift = (unsigned)((j + 1) * P.basebit);
          double p = ((double)k * (double)s1[i]) / (double)(1ull << shift);
          lwe_encrypt_f64(P, p, P.alpha_lv0, s0, rng, row);
        }
    });
  }
  if (bsk_fft) {
    std::vector<PolyEval*> evs(threads < 1 ? 1 : threads);
    for (auto& e : evs) e = new PolyEval(P.N);
    const size_t row = (size_t)2 * P.L * 2 * P.N;
    parallel_for(P.n, threads, [&](int i, int tid) {                           // :123-145
      Rng rng(seed * 0x100000001B3ull + 0x2000000ull + (uint64_t)i);
      trgsw_encrypt_fft(*evs[tid], P, s0[i], P.alpha_lv1, s1, rng, bsk_fft + row * i);
    });
    for (auto& e : evs) delete e;
  }
}

// single-step entry points for kernel-level parity tests
void oracle_external_product(void* e, const oracle_params* P, const double* bsk_row, const uint32_t* in /*[2][N]*/,
                             uint32_t offset, uint32_t* out /*[2][N]*/) {
  external_product(*(PolyEval*)e, *P, bsk_row, in, in + P->N, offset, out, out + P->N);
}
void oracle_cmux(void* e, const oracle_params* P, const double* bsk_row, const uint32_t* ct0, const uint32_t* ct1,
                 uint32_t offset, uint32_t* out) {
  cmux(*(PolyEval*)e, *P, bsk_row, ct0, ct0 + P->N, ct1, ct1 + P->N, offset, out, out + P->N);
}
void oracle_blind_rotate(void* e, const oracle_params* P, const uint32_t* ct, const uint32_t* testvec,
                         const double* bsk, uint32_t offset, uint32_t* out) {
  blind_rotate(*(PolyEval*)e, *P, ct, testvec, bsk, offset, out);
}
void oracle_sample_extract0(const uint32_t* trlwe, int N, uint32_t* out) { sample_extract0(trlwe, N, out); }
void oracle_key_switch(const oracle_params* P, const uint32_t* src, const uint32_t* ksk, uint32_t* out) {
  key_switch(*P, src, ksk, out);
}

// proxyreenc.ReencryptTLWELv0 (proxyreenc/proxyreenc.go:321-366).  key: KeyEncryptions flattened [n*t*base][n+1] with row
// idx = base*t*i + base*j + k (proxyreenc.go:286); base = 2^basebit and t are the key's own (ProxyReencryptionKey.Base, .T).
void oracle_reencrypt(const oracle_params* P, const uint32_t* ct_from, const uint32_t* key, int basebit, int t, uint32_t* out) {
  const int n = P->n, base = 1 << basebit;
  for (int x = 0; x <= n; x++) out[x] = 0;
  out[n] = ct_from[n];                                              // :338-339 result.SetB(ctFrom.B())
  const Torus prec = (Torus)1 << (32 - (1 + basebit * t));          // :342
  for (int i = 0; i < n; i++) {                                     // :345
    const Torus abar = ct_from[i] + prec;                           // :347
    for (int j = 0; j < t; j++) {                                   // :350
      const int shift = 32 - (j + 1) * basebit;                     // :352
      const Torus mask = ((Torus)1 << basebit) - 1;                 // :353
      const Torus k = (abar >> shift) & mask;                       // :354
      if (k != 0) {                                                 // :356
        const size_t idx = (size_t)base * t * i + (size_t)base * j + (size_t)k;  // :358
        const Torus* row = key + idx * (n + 1);
        for (int x = 0; x <= n; x++) out[x] = out[x] - row[x];      // :361-363
      }
    }
  }
}
int oracle_gate_prepare(const oracle_params* P, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  return gate_prepare(*P, op, a, b, out);
}

// Batch of independent bootstraps: the *intended* semantics of trgsw.BatchBlindRotate +
// gates.Batch* (trgsw/trgsw.go:234-252, gates/gates.go:156-312) — every element equals the
// single-gate path — run with one worker thread per requested core, each with private scratch
// (the sync.Pool evaluator equivalent).  luts: NULL => testvec for all; else [nluts][2][N] with
// lut index g % nluts... no: lut_index[g] selects (NULL => g if nluts == count else 0).
void oracle_bootstrap_batch(const oracle_params* Pp, int count, const uint32_t* ct_in, const uint32_t* testvec,
                            const uint32_t* luts, int nluts, const double* bsk, const uint32_t* ksk, uint32_t offset,
                            int threads, uint32_t* ct_out) {
  const oracle_params P = *Pp;
  if (threads < 1) threads = 1;
  std::vector<PolyEval*> evs(threads);
  for (auto& e : evs) e = new PolyEval(P.N);
  parallel_for(count, threads, [&](int g, int tid) {
    const Torus* tv = testvec;
    if (luts) tv = luts + (size_t)(nluts == count ? g : (nluts == 1 ? 0 : g % nluts)) * 2 * P.N;
    bootstrap(*evs[tid], P, ct_in + (size_t)g * (P.n + 1), tv, bsk, ksk, offset, ct_out + (size_t)g * (P.n + 1));
  });
  for (auto& e : evs) delete e;
}

// Batch of two-input gates: prologue + bootstrap with the default test vector.
// ops: one opcode per gate (or a single opcode when nops == 1).
int oracle_gate_batch(const oracle_params* Pp, int count, const uint8_t* ops, int nops, const uint32_t* a,
                      const uint32_t* b, const uint32_t* testvec, const double* bsk, const uint32_t* ksk,
                      uint32_t offset, int threads, uint32_t* out) {
  const oracle_params P = *Pp;
  std::vector<Torus> prep((size_t)count * (P.n + 1));
  for (int g = 0; g < count; g++) {
    int op = ops[nops == 1 ? 0 : g];
    if (gate_prepare(P, op, a + (size_t)g * (P.n + 1), b + (size_t)g * (P.n + 1), prep.data() + (size_t)g * (P.n + 1)))
      return -1;
  }
  oracle_bootstrap_batch(Pp, count, prep.data(), testvec, nullptr, 0, bsk, ksk, offset, threads, out);
  return 0;
}

}  // extern "C"
