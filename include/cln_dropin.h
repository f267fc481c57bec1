/*
 * cln_dropin.h — the reference's own entry points for the verification path, re-implemented on top
 * of the batch engine (cln_sigverify.h).  Same names, argument meaning and error behaviour as CLN:
 *
 *   check_signed_hash          bitcoin/signature.h:85    (bitcoin/signature.c:174-192)
 *   check_signed_hash_nodeid   common/node_id.h:80       (common/node_id.c:72-80)
 *   check_schnorr_sig          bitcoin/signature.h:129   (bitcoin/signature.c:408-430)
 *   sha256_double              bitcoin/shadouble.h       (bitcoin/shadouble.c:7-11)
 *   pubkey_from_der            bitcoin/pubkey.h          (bitcoin/pubkey.c:14-24)
 *   sigcheck_channel_announcement / _node_announcement / _channel_update
 *                              gossipd/sigcheck.h        (gossipd/sigcheck.c:45-115, 118-164, 9-43)
 *                              — here in BATCH form: n raw wire messages in, one status per message out
 *   check_tx_sigs_batch        the per-HTLC loop of channeld/channeld.c:2215-2232 (one shared key,
 *                              n sighashes, n signatures) as one launch
 *
 * check_tx_sig itself (bitcoin/signature.c:194-221) needs no replacement: it computes the BIP143
 * sighash with libwally on the host and then calls check_signed_hash — it picks up this
 * implementation unchanged.
 *
 * A false return always means "signature invalid" (peer's fault).  Engine failures (no GPU, CUDA
 * error) abort() with a message on stderr, CLN's convention for internal errors
 * (bitcoin/signature.c:117,212,420) — they are never reported as false.
 *
 * Types: inside CLN the real headers provide these (define CLN_TYPES_PROVIDED before including);
 * stand-alone, layout-compatible minimal definitions are supplied below.
 */
#ifndef CLN_DROPIN_H
#define CLN_DROPIN_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CLN_TYPES_PROVIDED
typedef unsigned char u8;
struct sha256 { union { uint32_t u32[8]; unsigned char u8[32]; } u; };  /* ccan/crypto/sha256/sha256.h */
struct sha256_double { struct sha256 sha; };                           /* bitcoin/shadouble.h:9-11 */
typedef struct { unsigned char data[64]; } secp256k1_ecdsa_signature;   /* secp256k1.h (opaque) */
typedef struct { unsigned char data[64]; } secp256k1_pubkey;            /* secp256k1.h (opaque) */
struct pubkey { secp256k1_pubkey pubkey; };                            /* bitcoin/pubkey.h:15-18 */
struct node_id { u8 k[33]; };                                          /* common/node_id.h:11-13 */
struct bip340sig { u8 u8[64]; };                                       /* bitcoin/signature.h:145-147 */
enum sighash_type { SIGHASH_ALL = 1, SIGHASH_NONE = 2, SIGHASH_SINGLE = 3, SIGHASH_ANYONECANPAY = 0x80 };
struct bitcoin_signature { secp256k1_ecdsa_signature s; enum sighash_type sighash_type; }; /* signature.h:47-50 */
#endif

/* Optional: choose the CUDA device (default: $CLN_SIGVERIFY_DEVICE or 0).  The context is created
 * lazily on first use; one per process, as CLN's global secp256k1_ctx (common/utils.c:16). */
void cln_sigverify_init(int device);
void cln_sigverify_shutdown(void);

bool check_signed_hash(const struct sha256_double *hash, const secp256k1_ecdsa_signature *signature,
                       const struct pubkey *key);
bool check_signed_hash_nodeid(const struct sha256_double *hash, const secp256k1_ecdsa_signature *signature,
                              const struct node_id *id);
bool check_schnorr_sig(const struct sha256 *hash, const secp256k1_pubkey *pubkey, const struct bip340sig *sig);
void sha256_double(struct sha256_double *shadouble, const void *p, size_t len);
bool pubkey_from_der(const u8 *der, size_t len, struct pubkey *key);

/* channeld HTLC loop: ok[i] = check_signed_hash(&hashes[i], &sigs[i].s, key) for one shared key. */
void check_tx_sigs_batch(const struct sha256_double *hashes, const struct bitcoin_signature *sigs,
                         const struct pubkey *key, size_t n, bool *ok);

/* The same loop with the BIP143 sighash ALSO computed on the device (row N2): check_tx_sig (bitcoin/signature.c:194-221)
 * for n one-input one-output transactions described by sv_tx records (cln_sigverify.h); the sighash type of each
 * signature is taken from sigs[i].sighash_type and gated exactly as check_tx_sig does (:206-211: only SIGHASH_ALL or
 * SIGHASH_SINGLE|SIGHASH_ANYONECANPAY, else false). */
struct sv_tx_fields; /* = sv_tx of cln_sigverify.h */
void check_tx_sigs_bip143_batch(const void *sv_tx_array, const u8 *scripts, size_t scripts_len,
                                const struct pubkey *key, const struct bitcoin_signature *sigs, size_t n, bool *ok);

/* gossipd: status[i] = 0 if every signature of message i verifies, else 1 + the index of the FIRST bad
 * signature in the reference's checking order (channel_announcement: node_signature_1, node_signature_2,
 * bitcoin_signature_1, bitcoin_signature_2 -> 1..4; node_announcement / channel_update: 1); -1 if the
 * message is too short / malformed to locate its fields.  msgs[i] is the complete wire message
 * (2-byte type included), lens[i] its length. */
void sigcheck_channel_announcement_batch(const u8 *const *msgs, const size_t *lens, size_t n, int *status);
void sigcheck_node_announcement_batch(const u8 *const *msgs, const size_t *lens, size_t n, int *status);
/* channel_update is signed by the node found in the gossmap: the caller supplies it. */
void sigcheck_channel_update_batch(const u8 *const *msgs, const size_t *lens, const struct node_id *signers,
                                   size_t n, int *status);

#ifdef __cplusplus
}
#endif
#endif
