// Reads like gates/gates_test.go: fresh keys, every input pair, assert on the decrypted Boolean.
// Built by tests/test_host_mirror.py with g++ against the C ABI; needs a GPU to run.
#include <cstdio>
#include "../../go-tfhe_b200/host/gotfhe.hpp"
using namespace gotfhe;

static int failures = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { failures++; std::printf("FAIL: " __VA_ARGS__); std::printf("\n"); } } while (0)

int main(int argc, char** argv) {
  const params::Set& P = params::get(argc > 1 ? argv[1] : "80");
  auto sk = key::NewSecretKey(P, 42);
  auto ck = cloudkey::NewCloudKey(sk, 43);
  struct { const char* name; gates::Ciphertext (*f)(const gates::Ciphertext&, const gates::Ciphertext&, cloudkey::CloudKey&); bool t[4]; } tab[] = {
      {"NAND", gates::NAND, {true, true, true, false}}, {"AND", gates::AND, {false, false, false, true}},
      {"OR", gates::OR, {false, true, true, true}},     {"XOR", gates::XOR, {false, true, true, false}},
      {"XNOR", gates::XNOR, {true, false, false, true}}, {"NOR", gates::NOR, {true, false, false, false}},
      {"ANDNY", gates::ANDNY, {false, true, false, false}}, {"ANDYN", gates::ANDYN, {false, false, true, false}},
      {"ORNY", gates::ORNY, {true, true, false, true}},  {"ORYN", gates::ORYN, {true, false, true, true}}};
  uint64_t seed = 100;
  for (auto& g : tab)
    for (int k = 0; k < 4; k++) {
      bool a = k >> 1, b = k & 1;
      auto r = g.f(tlwe::EncryptBool(a, sk, seed++), tlwe::EncryptBool(b, sk, seed++), *ck);
      EXPECT(tlwe::DecryptBool(r, sk) == g.t[k], "%s(%d,%d)", g.name, a, b);
    }
  for (int k = 0; k < 8; k++) {  // gates_test.go:338-366
    bool a = k >> 2, b = (k >> 1) & 1, c = k & 1;
    auto r = gates::MUX(tlwe::EncryptBool(a, sk, seed++), tlwe::EncryptBool(b, sk, seed++), tlwe::EncryptBool(c, sk, seed++), *ck);
    EXPECT(tlwe::DecryptBool(r, sk) == (a ? b : c), "MUX(%d,%d,%d)", a, b, c);
  }
  EXPECT(tlwe::DecryptBool(gates::NOT(tlwe::EncryptBool(true, sk, seed++)), sk) == false, "NOT");
  EXPECT(tlwe::DecryptBool(gates::Constant(true, P), sk) == true && tlwe::DecryptBool(gates::Constant(false, P), sk) == false, "Constant");
  evaluator::Evaluator ev(*ck);  // evaluator/programmable_bootstrap_test.go:13-105
  for (int m = 0; m < 2; m++) {
    auto ct = tlwe::EncryptLWEMessage(m, 2, sk, seed++);
    EXPECT(tlwe::DecryptLWEMessage(ev.BootstrapFunc(ct, [](int x) { return x; }, 2), 2, sk) == m, "identity(%d)", m);
    EXPECT(tlwe::DecryptLWEMessage(ev.BootstrapFunc(ct, [](int x) { return 1 - x; }, 2), 2, sk) == 1 - m, "not(%d)", m);
  }
  {  // the same truth table with a cloud key generated on the device
    auto dk = cloudkey::NewCloudKeyOnDevice(sk, 44);
    for (int k = 0; k < 4; k++) {
      bool a = k >> 1, b = k & 1;
      auto r = gates::NAND(tlwe::EncryptBool(a, sk, seed++), tlwe::EncryptBool(b, sk, seed++), *dk);
      EXPECT(tlwe::DecryptBool(r, sk) == !(a && b), "device key NAND(%d,%d)", a, b);
    }
    EXPECT(dk->DecompositionOffset == ck->DecompositionOffset && dk->BlindRotateTestvec == ck->BlindRotateTestvec, "device key constants");
  }
  {  // additive entry points: a full adder through the circuit runner (README.md:78-87) and a LUT table with indices
    std::vector<tfhe_gate_desc> fa = {{TFHE_OP_XOR, 0, 1, 0, 3}, {TFHE_OP_XOR, 3, 2, 0, 4}, {TFHE_OP_AND, 0, 1, 0, 5},
                                      {TFHE_OP_AND, 3, 2, 0, 6}, {TFHE_OP_OR, 5, 6, 0, 7}};
    std::vector<std::vector<gates::Ciphertext>> in(3);
    for (int k = 0; k < 8; k++)
      for (int w = 0; w < 3; w++) in[w].push_back(tlwe::EncryptBool((k >> w) & 1, sk, seed++));
    auto out = circuit::Run(fa, in, {4, 7}, *ck);
    for (int k = 0; k < 8; k++) {
      const int s = (k & 1) + ((k >> 1) & 1) + ((k >> 2) & 1);
      EXPECT(tlwe::DecryptBool(out[0][k], sk) == (bool)(s & 1) && tlwe::DecryptBool(out[1][k], sk) == (bool)(s >> 1), "full adder(%d)", k);
    }
    std::vector<lut::LookUpTable> luts = {lut::GenLookUpTable(P, 2, [](int x) { return x; }), lut::GenLookUpTable(P, 2, [](int x) { return 1 - x; })};
    std::vector<tlwe::TLWELv0> cts;
    std::vector<int32_t> idx;
    for (int k = 0; k < 4; k++) { cts.push_back(tlwe::EncryptLWEMessage(k & 1, 2, sk, seed++)); idx.push_back(k >> 1); }
    auto r = ev.BootstrapBatchIndexed(cts, luts, idx);
    for (int k = 0; k < 4; k++) EXPECT(tlwe::DecryptLWEMessage(r[k], 2, sk) == ((k >> 1) ? 1 - (k & 1) : (k & 1)), "indexed LUT(%d)", k);
  }
  std::printf("%s: %d failures\n", P.name, failures);
  return failures ? 1 : 0;
}
