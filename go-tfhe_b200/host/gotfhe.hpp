// gotfhe.hpp — header-only C++ mirror of the reference's Go packages for the bootstrap path, over the C ABI
// (include/tfhe_b200.h for the GPU engine, include/tfhe_b200_client.h for the host-side client helpers).
// Same names and argument meaning as the reference; errors throw std::runtime_error (the reference panics).
//   params::     params/params.go            key::       key/key.go
//   tlwe::       tlwe/tlwe.go, programmable_encrypt.go   cloudkey::  cloudkey/cloudkey.go
//   lut::        lut/generator.go            evaluator:: evaluator/evaluator.go, programmable_bootstrap.go
//   gates::      gates/gates.go
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "../../include/tfhe_b200_client.h"

namespace gotfhe {

using Torus = uint32_t;  // params/params.go:27

namespace params {
struct Set {  // params.TLWELv0Params + TRGSWLv1Params (params/params.go:50-78)
  const char* name; int n; double alpha_lv0; int N; double alpha_lv1; int NBIT, BGBIT, L, BASEBIT, IKS_T;
  tfhe_params c() const { return tfhe_params{n, N, L, BGBIT, BASEBIT, IKS_T}; }
  int ksk_rows() const { return N * IKS_T * (1 << BASEBIT); }
};
inline const Set& get(const std::string& name = "128") {  // params/params.go:83-391
  static const Set sets[] = {
      {"80", 550, 5.0e-5, 1024, 3.73e-8, 10, 6, 3, 2, 7},
      {"110", 630, 3.0517578125e-05, 1024, 2.980232238769531e-8, 10, 6, 3, 2, 8},
      {"128", 700, 2.0e-5, 1024, 2.0e-8, 10, 6, 3, 2, 9},
      {"uint1", 700, 2.0e-05, 1024, 2.0e-08, 10, 10, 2, 2, 8},
      {"uint2", 687, 0.00002120846893069971872305794214, 512, 0.00000000000231841227527049948463, 9, 18, 1, 4, 3},
      {"uint3", 820, 0.00000251676160959795544987084234, 1024, 0.00000000000000022204460492503131, 10, 23, 1, 6, 2},
      {"uint4", 820, 0.00000251676160959795544987084234, 2048, 0.00000000000000022204460492503131, 11, 22, 1, 5, 3},
      {"uint5", 1071, 7.088226765410429399593757e-08, 2048, 2.2204460492503131e-17, 11, 22, 1, 6, 3},
  };
  for (auto& s : sets) if (name == s.name) return s;
  throw std::runtime_error("unknown parameter set " + name);
}
}  // namespace params

namespace tlwe {
struct TLWELv0 { std::vector<Torus> P; };  // tlwe/tlwe.go:11-13 (length n+1, last element is b)
}
namespace trlwe {
struct TRLWELv1 { std::vector<Torus> A, B; };  // trlwe/trlwe.go:13-16
}

namespace key {
struct SecretKey { const params::Set* P; std::vector<Torus> KeyLv0, KeyLv1; };  // key/key.go:10-13
// seed == 0 (default): keyed from the operating system's entropy source, as the reference's unseeded generator; a non-zero
// seed makes the result reproducible (tests).  The same convention holds for every seed argument below.
inline SecretKey NewSecretKey(const params::Set& P, uint64_t seed = 0) {  // key/key.go:16-45
  SecretKey sk{&P, std::vector<Torus>(P.n), std::vector<Torus>(P.N)};
  tfhe_params c = P.c();
  tfhe_client_secret_key(&c, seed, sk.KeyLv0.data(), sk.KeyLv1.data());
  return sk;
}
}  // namespace key

namespace tlwe {
inline TLWELv0 EncryptBool(bool b, const key::SecretKey& sk, uint64_t seed) {  // tlwe/tlwe.go:54-62
  TLWELv0 ct{std::vector<Torus>(sk.P->n + 1)};
  tfhe_params c = sk.P->c();
  uint8_t bit = b;
  tfhe_client_encrypt_bool(&c, sk.P->alpha_lv0, sk.KeyLv0.data(), seed, 1, &bit, ct.P.data());
  return ct;
}
inline bool DecryptBool(const TLWELv0& ct, const key::SecretKey& sk) {  // tlwe/tlwe.go:65-74
  tfhe_params c = sk.P->c();
  uint8_t bit = 0;
  tfhe_client_decrypt_bool(&c, sk.KeyLv0.data(), 1, ct.P.data(), &bit);
  return bit != 0;
}
inline TLWELv0 EncryptLWEMessage(int m, int modulus, const key::SecretKey& sk, uint64_t seed) {  // programmable_encrypt.go:12-27
  TLWELv0 ct{std::vector<Torus>(sk.P->n + 1)};
  tfhe_params c = sk.P->c();
  int32_t mm = m;
  tfhe_client_encrypt_message(&c, sk.P->alpha_lv0, sk.KeyLv0.data(), seed, 1, &mm, modulus, ct.P.data());
  return ct;
}
inline int DecryptLWEMessage(const TLWELv0& ct, int modulus, const key::SecretKey& sk) {  // programmable_encrypt.go:33-54
  tfhe_params c = sk.P->c();
  int32_t m = 0;
  tfhe_client_decrypt_message(&c, sk.KeyLv0.data(), 1, ct.P.data(), modulus, &m);
  return m;
}
}  // namespace tlwe

namespace cloudkey {
// cloudkey.CloudKey (cloudkey/cloudkey.go:16-21), flattened; owns the GPU context it is loaded on.
struct CloudKey {
  const params::Set* P = nullptr;
  Torus DecompositionOffset = 0;
  std::vector<Torus> BlindRotateTestvec, KeySwitchingKey;
  std::vector<double> BootstrappingKey;
  tfhe_ctx* ctx = nullptr;
  CloudKey() = default;
  CloudKey(const CloudKey&) = delete;
  CloudKey& operator=(const CloudKey&) = delete;
  ~CloudKey() { if (ctx) tfhe_ctx_destroy(ctx); }
  std::vector<int> devices;  // GPUs the engine uses; empty = every visible GPU (tfhe_ctx_create_multi)
  tfhe_ctx* engine() {  // created and uploaded on first use
    if (!ctx) {
      tfhe_params c = P->c();
      if (tfhe_ctx_create_multi(&c, (int)devices.size(), devices.empty() ? nullptr : devices.data(), &ctx) != 0)
        throw std::runtime_error(std::string("tfhe_ctx_create_multi: ") + tfhe_last_error(nullptr));
      if (tfhe_ctx_load_cloudkey(ctx, DecompositionOffset, BootstrappingKey.data(), KeySwitchingKey.data(), BlindRotateTestvec.data()) != 0)
        throw std::runtime_error(std::string("tfhe_ctx_load_cloudkey: ") + tfhe_last_error(ctx));
    }
    return ctx;
  }
};
inline std::unique_ptr<CloudKey> NewCloudKey(const key::SecretKey& sk, uint64_t seed = 0) {  // cloudkey.go:24-31
  auto ck = std::make_unique<CloudKey>();
  const params::Set& P = *sk.P;
  ck->P = &P;
  ck->BlindRotateTestvec.resize(2 * P.N);
  ck->KeySwitchingKey.resize((size_t)P.ksk_rows() * (P.n + 1));
  ck->BootstrappingKey.resize((size_t)P.n * 2 * P.L * 2 * P.N);
  tfhe_params c = P.c();
  tfhe_client_cloud_key(&c, P.alpha_lv0, P.alpha_lv1, sk.KeyLv0.data(), sk.KeyLv1.data(), seed, 0, &ck->DecompositionOffset,
                        ck->BlindRotateTestvec.data(), ck->KeySwitchingKey.data(), ck->BootstrappingKey.data());
  return ck;
}
// cloudkey.NewCloudKey with the key material generated on the device (tfhe_ctx_generate_cloudkey): same fields, and the
// engine that made them already holds the key, so no second upload happens.
inline std::unique_ptr<CloudKey> NewCloudKeyOnDevice(const key::SecretKey& sk, uint64_t seed = 0, const std::vector<int>& devices = {}) {
  auto ck = std::make_unique<CloudKey>();
  const params::Set& P = *sk.P;
  ck->P = &P;
  ck->BlindRotateTestvec.resize(2 * P.N);
  ck->KeySwitchingKey.resize((size_t)P.ksk_rows() * (P.n + 1));
  ck->BootstrappingKey.resize((size_t)P.n * 2 * P.L * 2 * P.N);
  tfhe_params c = P.c();
  if (tfhe_ctx_create_multi(&c, (int)devices.size(), devices.empty() ? nullptr : devices.data(), &ck->ctx) != 0)
    throw std::runtime_error(std::string("tfhe_ctx_create_multi: ") + tfhe_last_error(nullptr));
  if (tfhe_ctx_generate_cloudkey(ck->ctx, sk.KeyLv0.data(), sk.KeyLv1.data(), P.alpha_lv0, P.alpha_lv1, seed, 1, &ck->DecompositionOffset,
                                 ck->BootstrappingKey.data(), ck->KeySwitchingKey.data(), ck->BlindRotateTestvec.data()) != 0)
    throw std::runtime_error(std::string("tfhe_ctx_generate_cloudkey: ") + tfhe_last_error(ck->ctx));
  return ck;
}
}  // namespace cloudkey

namespace lut {
struct LookUpTable { trlwe::TRLWELv1 Poly; };  // lut/lut.go:14-17
inline LookUpTable GenLookUpTable(const params::Set& P, int messageModulus, const std::function<int(int)>& f) {  // generator.go:49-100
  std::vector<int32_t> fv(messageModulus);
  for (int x = 0; x < messageModulus; x++) fv[x] = f(x);
  std::vector<Torus> flat(2 * P.N);
  tfhe_params c = P.c();
  tfhe_client_gen_lut(&c, messageModulus, fv.data(), flat.data());
  LookUpTable t;
  t.Poly.A.assign(flat.begin(), flat.begin() + P.N);
  t.Poly.B.assign(flat.begin() + P.N, flat.end());
  return t;
}
}  // namespace lut

namespace detail {
inline void check(tfhe_ctx* ctx, int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + tfhe_last_error(ctx));
}
inline std::vector<Torus> flatten(const std::vector<tlwe::TLWELv0>& v, int n1) {
  std::vector<Torus> out;
  out.reserve(v.size() * n1);
  for (auto& c : v) out.insert(out.end(), c.P.begin(), c.P.end());
  return out;
}
inline std::vector<tlwe::TLWELv0> unflatten(const std::vector<Torus>& buf, size_t count, int n1) {
  std::vector<tlwe::TLWELv0> out(count);
  for (size_t i = 0; i < count; i++) out[i].P.assign(buf.begin() + i * n1, buf.begin() + (i + 1) * n1);
  return out;
}
}  // namespace detail

namespace evaluator {
// evaluator.Evaluator (evaluator/evaluator.go:15-35): scratch lives in the CloudKey's GPU context.
struct Evaluator {
  cloudkey::CloudKey* ck;
  explicit Evaluator(cloudkey::CloudKey& k) : ck(&k) {}
  // Bootstrap / BootstrapAssign (evaluator.go:139-157) and BootstrapLUT (programmable_bootstrap.go:54-115), batched
  std::vector<tlwe::TLWELv0> BootstrapBatch(const std::vector<tlwe::TLWELv0>& cts, const lut::LookUpTable* lut = nullptr) {
    const int n1 = ck->P->n + 1;
    auto in = detail::flatten(cts, n1);
    std::vector<Torus> out(in.size()), lflat;
    if (lut) { lflat = lut->Poly.A; lflat.insert(lflat.end(), lut->Poly.B.begin(), lut->Poly.B.end()); }
    detail::check(ck->engine(), tfhe_bootstrap_batch(ck->engine(), (int64_t)cts.size(), in.data(), lut ? lflat.data() : nullptr,
                                                     lut ? 1 : 0, out.data()), "tfhe_bootstrap_batch");
    return detail::unflatten(out, cts.size(), n1);
  }
  // additive: a LUT table shared by the batch, ciphertext k uses luts[index[k]] (tfhe_bootstrap_batch_indexed)
  std::vector<tlwe::TLWELv0> BootstrapBatchIndexed(const std::vector<tlwe::TLWELv0>& cts, const std::vector<lut::LookUpTable>& luts,
                                                   const std::vector<int32_t>& index) {
    const int n1 = ck->P->n + 1;
    auto in = detail::flatten(cts, n1);
    std::vector<Torus> out(in.size()), lflat;
    for (auto& l : luts) { lflat.insert(lflat.end(), l.Poly.A.begin(), l.Poly.A.end()); lflat.insert(lflat.end(), l.Poly.B.begin(), l.Poly.B.end()); }
    detail::check(ck->engine(), tfhe_bootstrap_batch_indexed(ck->engine(), (int64_t)cts.size(), in.data(), lflat.data(), (int64_t)luts.size(),
                                                             index.data(), out.data()), "tfhe_bootstrap_batch_indexed");
    return detail::unflatten(out, cts.size(), n1);
  }
  tlwe::TLWELv0 Bootstrap(const tlwe::TLWELv0& ct) { return BootstrapBatch({ct})[0]; }
  tlwe::TLWELv0 BootstrapLUT(const tlwe::TLWELv0& ct, const lut::LookUpTable& l) { return BootstrapBatch({ct}, &l)[0]; }
  tlwe::TLWELv0 BootstrapFunc(const tlwe::TLWELv0& ct, const std::function<int(int)>& f, int messageModulus) {  // :16-29
    return BootstrapLUT(ct, lut::GenLookUpTable(*ck->P, messageModulus, f));
  }
};
}  // namespace evaluator

namespace gates {
using Ciphertext = tlwe::TLWELv0;  // gates/gates.go:16
inline std::vector<Ciphertext> Batch(tfhe_op op, const std::vector<Ciphertext>& a, const std::vector<Ciphertext>& b,
                                     cloudkey::CloudKey& ck, const std::vector<Ciphertext>* c = nullptr) {
  const int n1 = ck.P->n + 1;
  auto fa = detail::flatten(a, n1), fb = detail::flatten(b, n1);
  std::vector<Torus> fc, out(fa.size());
  if (c) fc = detail::flatten(*c, n1);
  uint8_t o = (uint8_t)op;
  detail::check(ck.engine(), tfhe_gate_batch(ck.engine(), (int64_t)a.size(), &o, 1, fa.data(), fb.data(), c ? fc.data() : nullptr, out.data()),
                "tfhe_gate_batch");
  return detail::unflatten(out, a.size(), n1);
}
#define GOTFHE_GATE(NAME, OP) \
  inline Ciphertext NAME(const Ciphertext& a, const Ciphertext& b, cloudkey::CloudKey& ck) { return Batch(OP, {a}, {b}, ck)[0]; }
GOTFHE_GATE(NAND, TFHE_OP_NAND)    // gates.go:26
GOTFHE_GATE(OR, TFHE_OP_OR)        // :34
GOTFHE_GATE(AND, TFHE_OP_AND)      // :40
GOTFHE_GATE(XOR, TFHE_OP_XOR)      // :46
GOTFHE_GATE(XNOR, TFHE_OP_XNOR)    // :52
GOTFHE_GATE(NOR, TFHE_OP_NOR)      // :72
GOTFHE_GATE(ANDNY, TFHE_OP_ANDNY)  // :79
GOTFHE_GATE(ANDYN, TFHE_OP_ANDYN)  // :86
GOTFHE_GATE(ORNY, TFHE_OP_ORNY)    // :93
GOTFHE_GATE(ORYN, TFHE_OP_ORYN)    // :100
#undef GOTFHE_GATE
inline Ciphertext MUX(const Ciphertext& a, const Ciphertext& b, const Ciphertext& c, cloudkey::CloudKey& ck) {  // :107-114
  std::vector<Ciphertext> cc{c};
  return Batch(TFHE_OP_MUX, {a}, {b}, ck, &cc)[0];
}
inline Ciphertext NOT(const Ciphertext& a) {  // :117-119
  Ciphertext r{a.P};
  for (auto& w : r.P) w = 0u - w;
  return r;
}
inline Ciphertext Copy(const Ciphertext& a) { return a; }  // :122-126
inline Ciphertext Constant(bool v, const params::Set& P) {  // :61-69
  Ciphertext r{std::vector<Torus>(P.n + 1, 0)};
  r.P[P.n] = v ? 0x20000000u : (Torus)(1u - 0x20000000u);
  return r;
}
}  // namespace gates

namespace circuit {
// additive: the same circuit on many instances, one batch per level, wires resident on the GPUs.
// inputs[w][k] = input wire w of instance k; the result is indexed [output][instance] (tfhe_circuit_run).
inline std::vector<std::vector<gates::Ciphertext>> Run(const std::vector<tfhe_gate_desc>& gate_list,
                                                       const std::vector<std::vector<gates::Ciphertext>>& inputs,
                                                       const std::vector<int32_t>& output_wires, cloudkey::CloudKey& ck) {
  const int n1 = ck.P->n + 1;
  const size_t instances = inputs.empty() ? 0 : inputs[0].size();
  std::vector<Torus> in, out(output_wires.size() * instances * n1);
  for (auto& w : inputs) { auto f = detail::flatten(w, n1); in.insert(in.end(), f.begin(), f.end()); }
  detail::check(ck.engine(), tfhe_circuit_run(ck.engine(), (int64_t)instances, (int32_t)inputs.size(), (int32_t)gate_list.size(), gate_list.data(),
                                              in.data(), (int32_t)output_wires.size(), output_wires.data(), out.data()), "tfhe_circuit_run");
  std::vector<std::vector<gates::Ciphertext>> res(output_wires.size());
  for (size_t k = 0; k < res.size(); k++)
    res[k] = detail::unflatten(std::vector<Torus>(out.begin() + k * instances * n1, out.begin() + (k + 1) * instances * n1), instances, n1);
  return res;
}
}  // namespace circuit
}  // namespace gotfhe
