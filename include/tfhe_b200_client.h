/*
 * tfhe_b200_client.h — host-side client helpers (libtfhe_b200_client.so, plain C++, no CUDA).
 *
 * NOT part of the drop-in boundary: in an integration with the reference these operations stay
 * in its own Go packages (key/key.go, tlwe/tlwe.go, tlwe/programmable_encrypt.go,
 * cloudkey/cloudkey.go, lut/generator.go).  They exist here so that bench.py, the examples and the
 * host-side mirror can make valid keys and ciphertexts on a machine without a Go toolchain.
 * Layouts are those of tfhe_b200.h.
 */
#ifndef TFHE_B200_CLIENT_H
#define TFHE_B200_CLIENT_H
#include "tfhe_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* --- TFHB wire format: keys and ciphertext batches as flat files (the reference has no serialisation at all; its
 * "format" is the Go structs key.SecretKey key/key.go:10-13 and cloudkey.CloudKey cloudkey/cloudkey.go:16-21).  The same
 * format is written by go/tfheb200/wire.go (so that a real Go run can ship keys and golden vectors) and by
 * go-tfhe_b200/wire.py.  Layout: "TFHB", version u32 = 2, kind u32, params 6 x i32, nsect u32, sections (tag[4], dtype
 * u32: 0 = u32 / 1 = f64, count u64, data), CRC-32 (IEEE) of everything before it as u64.  Kinds: 1 SecretKey (lv0, lv1),
 * 2 CloudKey (offs, tvec, bsk, [ksk]), 3 ciphertext batch (ct), 4 TRLWE / LUT batch (trlw), 5 named vector bundle. */
typedef struct { char tag[4]; uint32_t dtype; uint64_t count; const void* data; } tfhe_wire_section;
/* out == NULL: size needed; else bytes written, or -1 (buffer too small / bad argument) */
int64_t tfhe_wire_pack(uint32_t kind, const tfhe_params* P, const tfhe_wire_section* sections, uint32_t nsect, void* out,
                       int64_t out_cap);
/* 0 ok (section data pointers point into blob); -1 malformed, -2 checksum, -3 version, -4 *nsect (capacity) too small;
 * on return *nsect = sections in the blob */
int tfhe_wire_unpack(const void* blob, int64_t size, uint32_t* kind, tfhe_params* P, tfhe_wire_section* sections,
                     uint32_t* nsect);
/* Randomness: every `seed` below selects the ChaCha20 key of that call — 0 = 256 bits from the OS entropy source
 * (what a real client uses; the Python mirror's default), non-zero = reproducible expansion of the seed (tests). */
/* ChaCha20 block function (RFC 8439) used by this library and by the device key generator; for known-answer tests */
void tfhe_client_chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]);
/* key.NewSecretKey, key/key.go:16-45 */
void tfhe_client_secret_key(const tfhe_params* P, uint64_t seed, uint32_t* key_lv0, uint32_t* key_lv1);
/* tlwe.EncryptBool / DecryptBool, tlwe/tlwe.go:54-74 */
void tfhe_client_encrypt_bool(const tfhe_params* P, double alpha, const uint32_t* key_lv0, uint64_t seed,
                              int64_t count, const uint8_t* bits, uint32_t* out);
void tfhe_client_decrypt_bool(const tfhe_params* P, const uint32_t* key_lv0, int64_t count, const uint32_t* ct,
                              uint8_t* bits);
/* tlwe.EncryptLWEMessage / DecryptLWEMessage, tlwe/programmable_encrypt.go:12-54 */
void tfhe_client_encrypt_message(const tfhe_params* P, double alpha, const uint32_t* key_lv0, uint64_t seed,
                                 int64_t count, const int32_t* msgs, int32_t modulus, uint32_t* out);
void tfhe_client_decrypt_message(const tfhe_params* P, const uint32_t* key_lv0, int64_t count, const uint32_t* ct,
                                 int32_t modulus, int32_t* msgs);
/* lut.Generator.GenLookUpTable, lut/generator.go:49-100; fvals[x] = f(x); lut_out is a TRLWE [2][N] */
void tfhe_client_gen_lut(const tfhe_params* P, int32_t modulus, const int32_t* fvals, uint32_t* lut_out);
/* cloudkey.NewCloudKey, cloudkey/cloudkey.go:24-31,60-145 */
void tfhe_client_cloud_key(const tfhe_params* P, double alpha_lv0, double alpha_lv1, const uint32_t* key_lv0,
                           const uint32_t* key_lv1, uint64_t seed, int threads, uint32_t* decomposition_offset,
                           uint32_t* testvec, uint32_t* ksk, double* bsk_fft);
#ifdef __cplusplus
}
#endif
#endif
