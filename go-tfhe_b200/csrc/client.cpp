// client.cpp — host-side client operations (key generation, encryption, decryption, LUT
// construction).  In a real integration these stay in the reference's own Go packages
// (key/, tlwe/, cloudkey/, lut/) and only the flattened results cross the C ABI; this
// library is the stand-in used where no Go toolchain exists (this image), so that bench.py
// and the examples can produce valid keys and ciphertexts WITHOUT touching oracle/.
// It is an independent implementation: integer-exact ring products (the secret keys are
// binary, so a*s is a signed sum of rotations), its own transform code, and ChaCha20 randomness
// (chacha.h) keyed from the OS entropy source unless a caller passes a non-zero seed for a
// reproducible run (tests, benchmarks).
//
// Produces exactly the structures the reference's CloudKey holds (cloudkey/cloudkey.go:16-21),
// flattened as documented in include/tfhe_b200.h.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/tfhe_b200_client.h"
#include "chacha.h"

namespace {

typedef uint32_t Torus;

// ChaCha20 streams (chacha.h): stream (domain, id) under the call's 256-bit key; results do not depend on the number of
// worker threads.  Masks and noise of one ciphertext come from different domains.
struct Stream {
  tfhe::RngKey key;
  uint32_t dom;
  uint64_t id;
  uint32_t buf[16], ctr = 0;
  int pos = 16;
  Stream(const tfhe::RngKey& k, uint32_t domain, uint64_t stream_id) : key(k), dom(domain), id(stream_id) {}
  uint32_t u32() {
    if (pos == 16) { tfhe::chacha20_block(key, ctr++, dom, (uint32_t)id, (uint32_t)(id >> 32), buf); pos = 0; }
    return buf[pos++];
  }
  double unit() { const uint32_t hi = u32(), lo = u32(); return tfhe::rng_unit(hi, lo); }
  double gauss() {  // Box-Muller, one value per call
    const double u = unit(), v = unit();
    return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586476925286766559 * v);
  }
};
tfhe::RngKey call_key(uint64_t seed) {  // seed == 0: OS entropy (the default of the Python mirror); else reproducible
  tfhe::RngKey k;
  if (!tfhe::rng_make_key(seed, &k)) { std::fprintf(stderr, "tfhe_b200_client: no entropy source\n"); std::abort(); }
  return k;
}

// real number -> torus, utils.F64ToTorus semantics (utils/utils.go:11-14): frac(d) * 2^32, truncated
Torus to_torus(double d) {
  double f = std::fmod(d, 1.0) * 4294967296.0;
  return (Torus)(uint64_t)(int64_t)f;
}
Torus noisy(double mu, double sigma, Stream& noise) { return to_torus(mu) + to_torus(noise.gauss() * sigma); }

void lwe_encrypt(const tfhe_params& P, double mu, double sigma, const Torus* s0, Stream& mask, Stream& noise, Torus* out) {
  Torus dot = 0;
  for (int i = 0; i < P.n; i++) {
    out[i] = mask.u32();
    dot += out[i] * s0[i];
  }
  out[P.n] = dot + noisy(mu, sigma, noise);
}

// (a * s) in Z[X]/(X^N+1) for binary s, exact mod 2^32
void ring_mul_binary(const Torus* a, const Torus* s, int N, Torus* out) {
  std::memset(out, 0, sizeof(Torus) * N);
  for (int j = 0; j < N; j++) {
    if (!s[j]) continue;
    for (int i = 0; i < N - j; i++) out[i + j] += a[i];
    for (int i = N - j; i < N; i++) out[i + j - N] -= a[i];
  }
}

// Transform of the reference's FourierPoly: evaluate the folded polynomial z_j = p_j + i p_{j+N/2}
// at the roots of x^(N/2) = i by repeatedly splitting x^m - c into x^(m/2) -+ sqrt(c).  The order of
// the outputs and the 4-real/4-imaginary packing are those of poly/fourier_transform.go.
struct Spectrum {
  int N, M;
  std::vector<double> wr, wi;  // sqrt table, stage-major: entry (m - 1 + i) for block i of stage m
  explicit Spectrum(int N_) : N(N_), M(N_ / 2), wr(M), wi(M) {
    int bits = 0;
    while ((1 << bits) < M / 2) bits++;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int m = 1; m <= M / 2; m <<= 1)
      for (int i = 0; i < m; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++)
          if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        long double ang = -2.0L * pi * r / M + pi / (4.0L * m);
        wr[m - 1 + i] = (double)cosl(ang);
        wi[m - 1 + i] = (double)sinl(ang);
      }
  }
  void forward(const Torus* p, double* out) const {
    std::vector<double> re(M), im(M);
    for (int j = 0; j < M; j++) { re[j] = (double)(int32_t)p[j]; im[j] = (double)(int32_t)p[j + M]; }
    for (int m = 1, half = M / 2; m <= M / 2; m <<= 1, half >>= 1)
      for (int i = 0; i < m; i++) {
        const double cr = wr[m - 1 + i], ci = wi[m - 1 + i];
        const int lo = 2 * i * half;
        for (int j = lo; j < lo + half; j++) {
          const double tr = re[j + half] * cr - im[j + half] * ci, ti = re[j + half] * ci + im[j + half] * cr;
          re[j + half] = re[j] - tr; im[j + half] = im[j] - ti;
          re[j] += tr; im[j] += ti;
        }
      }
    for (int k = 0; k < M; k++) { out[(k >> 2) * 8 + (k & 3)] = re[k]; out[(k >> 2) * 8 + 4 + (k & 3)] = im[k]; }
  }
};

template <class F> void run_parallel(int count, int threads, F f) {
  if (threads < 1) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads > count) threads = count > 0 ? count : 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([=]() { for (int i = t; i < count; i += threads) f(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// ---- TFHB wire format (go-tfhe_b200/wire.py, go/tfheb200/wire.go hold the same format) -----------------------------
//   "TFHB" | version u32 (2) | kind u32 | params 6 x i32 | nsect u32 | sections | crc u64
//   section: tag (4 ascii bytes, space padded) | dtype u32 (0 = u32, 1 = f64) | count u64 | count little-endian values
//   crc: CRC-32 (IEEE 802.3, the zlib / Go hash/crc32 polynomial) of every byte before it, zero-extended to 64 bits
static uint32_t wire_crc32(const unsigned char* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
static const uint32_t kWireVersion = 2;

// Serialises `nsect` sections.  out == NULL: returns the size needed.  Else writes at most out_cap bytes and returns the
// size written, or -1 if out_cap is too small / an argument is invalid.  Little-endian hosts only (x86-64, aarch64).
int64_t tfhe_wire_pack(uint32_t kind, const tfhe_params* P, const tfhe_wire_section* sections, uint32_t nsect, void* out,
                       int64_t out_cap) {
  if (!P || (nsect && !sections)) return -1;
  size_t need = 4 + 4 + 4 + 24 + 4 + 8;
  for (uint32_t i = 0; i < nsect; i++) {
    if (sections[i].dtype > 1 || (sections[i].count && !sections[i].data)) return -1;
    need += 16 + (size_t)sections[i].count * (sections[i].dtype ? 8 : 4);
  }
  if (!out) return (int64_t)need;
  if ((size_t)out_cap < need) return -1;
  unsigned char* w = static_cast<unsigned char*>(out);
  size_t off = 0;
  auto put = [&](const void* src, size_t n) { std::memcpy(w + off, src, n); off += n; };
  put("TFHB", 4);
  put(&kWireVersion, 4);
  put(&kind, 4);
  const int32_t pv[6] = {P->n, P->N, P->L, P->bgbit, P->basebit, P->iks_t};
  put(pv, 24);
  put(&nsect, 4);
  for (uint32_t i = 0; i < nsect; i++) {
    put(sections[i].tag, 4);
    put(&sections[i].dtype, 4);
    put(&sections[i].count, 8);
    put(sections[i].data, (size_t)sections[i].count * (sections[i].dtype ? 8 : 4));
  }
  const uint64_t crc = wire_crc32(w, off);
  put(&crc, 8);
  return (int64_t)off;
}

// Parses a blob: checks magic, version and checksum, fills kind / params and up to *nsect section descriptors whose
// `data` pointers point INTO the blob (valid while it is).  On return *nsect is the number of sections in the blob.
// Returns 0, or -1 malformed / truncated, -2 bad checksum, -3 unsupported version, -4 more sections than capacity.
int tfhe_wire_unpack(const void* blob, int64_t size, uint32_t* kind, tfhe_params* P, tfhe_wire_section* sections,
                     uint32_t* nsect) {
  if (!blob || !kind || !P || !nsect || size < 48) return -1;
  const unsigned char* r = static_cast<const unsigned char*>(blob);
  if (std::memcmp(r, "TFHB", 4) != 0) return -1;
  uint32_t version;
  std::memcpy(&version, r + 4, 4);
  if (version != kWireVersion) return -3;
  uint64_t crc;
  std::memcpy(&crc, r + size - 8, 8);
  if (crc != (uint64_t)wire_crc32(r, (size_t)size - 8)) return -2;
  std::memcpy(kind, r + 8, 4);
  int32_t pv[6];
  std::memcpy(pv, r + 12, 24);
  P->n = pv[0]; P->N = pv[1]; P->L = pv[2]; P->bgbit = pv[3]; P->basebit = pv[4]; P->iks_t = pv[5];
  uint32_t ns;
  std::memcpy(&ns, r + 36, 4);
  const uint32_t cap = *nsect;
  *nsect = ns;
  size_t off = 40;
  for (uint32_t i = 0; i < ns; i++) {
    if (off + 16 > (size_t)size - 8) return -1;
    tfhe_wire_section sec;
    std::memcpy(sec.tag, r + off, 4);
    std::memcpy(&sec.dtype, r + off + 4, 4);
    std::memcpy(&sec.count, r + off + 8, 8);
    off += 16;
    if (sec.dtype > 1) return -1;
    const size_t bytes = (size_t)sec.count * (sec.dtype ? 8 : 4);
    if (bytes > (size_t)size - 8 - off) return -1;
    sec.data = r + off;
    off += bytes;
    if (i < cap && sections) sections[i] = sec;
  }
  if (off != (size_t)size - 8) return -1;
  return ns > cap ? -4 : 0;
}

// The ChaCha20 block function of chacha.h (shared with the device key generator), exposed for the RFC 8439 known-answer test.
void tfhe_client_chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]) {
  tfhe::RngKey k;
  std::memcpy(k.k, key, sizeof k.k);
  uint32_t o[16];
  tfhe::chacha20_block(k, counter, nonce[0], nonce[1], nonce[2], o);
  std::memcpy(out, o, sizeof o);
}

// key.NewSecretKey (key/key.go:16-45): uniform binary keys of length n and N.
void tfhe_client_secret_key(const tfhe_params* P, uint64_t seed, uint32_t* key_lv0, uint32_t* key_lv1) {
  const tfhe::RngKey k = call_key(seed);
  Stream a(k, tfhe::RNG_SK_LV0, 0), b(k, tfhe::RNG_SK_LV1, 0);
  for (int i = 0; i < P->n; i++) key_lv0[i] = a.u32() >> 31;
  for (int i = 0; i < P->N; i++) key_lv1[i] = b.u32() >> 31;
}

// tlwe.EncryptBool (tlwe/tlwe.go:54-62): mu = +-1/8.  count ciphertexts, ciphertext g uses stream (seed, g).
void tfhe_client_encrypt_bool(const tfhe_params* P, double alpha, const uint32_t* key_lv0, uint64_t seed, int64_t count,
                              const uint8_t* bits, uint32_t* out) {
  const tfhe::RngKey k = call_key(seed);
  for (int64_t g = 0; g < count; g++) {
    Stream mask(k, tfhe::RNG_ENC_MASK, (uint64_t)g), noise(k, tfhe::RNG_ENC_NOISE, (uint64_t)g);
    lwe_encrypt(*P, bits[g] ? 0.125 : -0.125, alpha, key_lv0, mask, noise, out + (size_t)g * (P->n + 1));
  }
}
// tlwe.DecryptBool (tlwe/tlwe.go:65-74)
void tfhe_client_decrypt_bool(const tfhe_params* P, const uint32_t* key_lv0, int64_t count, const uint32_t* ct,
                              uint8_t* bits) {
  for (int64_t g = 0; g < count; g++) {
    const Torus* c = ct + (size_t)g * (P->n + 1);
    Torus dot = 0;
    for (int i = 0; i < P->n; i++) dot += c[i] * key_lv0[i];
    bits[g] = (int32_t)(c[P->n] - dot) >= 0;
  }
}
// tlwe.EncryptLWEMessage (tlwe/programmable_encrypt.go:12-27): mu = m / (2 * modulus)
void tfhe_client_encrypt_message(const tfhe_params* P, double alpha, const uint32_t* key_lv0, uint64_t seed,
                                 int64_t count, const int32_t* msgs, int32_t modulus, uint32_t* out) {
  const tfhe::RngKey k = call_key(seed);
  for (int64_t g = 0; g < count; g++) {
    Stream mask(k, tfhe::RNG_ENC_MASK, (1ull << 40) + (uint64_t)g), noise(k, tfhe::RNG_ENC_NOISE, (1ull << 40) + (uint64_t)g);
    int m = msgs[g] % modulus;
    if (m < 0) m += modulus;
    const double mu = (double)m * (2147483648.0 / (double)modulus) / 4294967296.0;
    lwe_encrypt(*P, mu, alpha, key_lv0, mask, noise, out + (size_t)g * (P->n + 1));
  }
}
// tlwe.DecryptLWEMessage (tlwe/programmable_encrypt.go:33-54)
void tfhe_client_decrypt_message(const tfhe_params* P, const uint32_t* key_lv0, int64_t count, const uint32_t* ct,
                                 int32_t modulus, int32_t* msgs) {
  const Torus scale = (Torus)2147483648u / (Torus)modulus;
  for (int64_t g = 0; g < count; g++) {
    const Torus* c = ct + (size_t)g * (P->n + 1);
    Torus dot = 0;
    for (int i = 0; i < P->n; i++) dot += c[i] * key_lv0[i];
    const Torus phase = c[P->n] - dot;
    msgs[g] = (int32_t)(((Torus)(phase + scale / 2) / scale) % (Torus)modulus);
  }
}

// lut.Generator.GenLookUpTable (lut/generator.go:49-100) for fvals[x] = f(x), x in [0, modulus):
// box x covers table slots [round(xN/mod), round((x+1)N/mod)), the table is rotated left by half a
// box and the wrapped tail is negated; output is a TRLWE (A = 0, B = table).
void tfhe_client_gen_lut(const tfhe_params* P, int32_t modulus, const int32_t* fvals, uint32_t* lut_out) {
  const int N = P->N;
  std::vector<Torus> box(N, 0);
  for (int x = 0; x < modulus; x++) {
    const long lo = ((long)x * N + modulus / 2) / modulus, hi = ((long)(x + 1) * N + modulus / 2) / modulus;
    int y = fvals[x] % modulus;
    if (y < 0) y += modulus;
    const Torus enc = to_torus((double)y * (1.0 / (double)(2 * modulus)));
    for (long k = lo; k < hi; k++) box[k] = enc;
  }
  const long half = ((long)N + modulus) / (2L * modulus);  // round(N / (2 mod))
  for (int i = 0; i < N; i++) {
    Torus v = box[(i + half) % N];
    lut_out[N + i] = (i >= N - half) ? (Torus)0 - v : v;
    lut_out[i] = 0;
  }
}

// cloudkey.NewCloudKey (cloudkey/cloudkey.go:24-31,60-145).  alpha_lv0 = KSKAlpha, alpha_lv1 = BSKAlpha.
// ksk / bsk_fft may be NULL to skip that part.  threads <= 0 => all hardware threads.
void tfhe_client_cloud_key(const tfhe_params* Pp, double alpha_lv0, double alpha_lv1, const uint32_t* key_lv0,
                           const uint32_t* key_lv1, uint64_t seed, int threads, uint32_t* decomposition_offset,
                           uint32_t* testvec, uint32_t* ksk, double* bsk_fft) {
  const tfhe_params P = *Pp;
  const tfhe::RngKey rk = call_key(seed);
  Torus off = 0;
  for (int l = 0; l < P.L; l++) off += (Torus)(1u << (P.bgbit - 1)) << (32 - (l + 1) * P.bgbit);
  *decomposition_offset = off;
  for (int i = 0; i < P.N; i++) { testvec[i] = 0; testvec[P.N + i] = 0x20000000u; }
  const int base = 1 << P.basebit;
  if (ksk)
    run_parallel(P.N, threads, [&](int i) {
      for (int j = 0; j < P.iks_t; j++)
        for (int k = 0; k < base; k++) {
          Torus* row = ksk + ((size_t)(i * P.iks_t + j) * base + k) * (P.n + 1);
          if (k == 0) { std::memset(row, 0, sizeof(Torus) * (P.n + 1)); continue; }
          const uint64_t id = (uint64_t)(i * P.iks_t + j) * base + k;
          Stream mask(rk, tfhe::RNG_KSK_MASK, id), noise(rk, tfhe::RNG_KSK_NOISE, id);
          const double mu = (double)k * (double)key_lv1[i] / (double)(1ull << ((j + 1) * P.basebit));
          lwe_encrypt(P, mu, alpha_lv0, key_lv0, mask, noise, row);
        }
    });
  if (bsk_fft) {
    const Spectrum spec(P.N);
    run_parallel(P.n, threads, [&](int i) {
      const int N = P.N;
      std::vector<Torus> A(N), B(N), as(N);
      for (int r = 0; r < 2 * P.L; r++) {  // TRGSW row r: TRLWE encryption of zero plus the gadget term
        const uint64_t id = (uint64_t)i * 2 * P.L + r;
        Stream mask(rk, tfhe::RNG_BSK_MASK, id), noise(rk, tfhe::RNG_BSK_NOISE, id);
        for (int k = 0; k < N; k++) A[k] = mask.u32();
        ring_mul_binary(A.data(), key_lv1, N, as.data());
        for (int k = 0; k < N; k++) B[k] = as[k] + noisy(0.0, alpha_lv1, noise);
        const int lvl = r % P.L;
        const Torus g = key_lv0[i] * ((Torus)1u << (32 - (lvl + 1) * P.bgbit));  // s_i / Bg^(lvl+1)
        if (r < P.L) A[0] += g; else B[0] += g;
        double* dst = bsk_fft + (((size_t)i * 2 * P.L + r) * 2) * N;
        spec.forward(A.data(), dst);
        spec.forward(B.data(), dst + N);
      }
    });
  }
}

}  // extern "C"
