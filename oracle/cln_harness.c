/*
 * oracle/cln_harness.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * BASELINE config C1 ("reference plumbing"): CLN's own, UNMODIFIED bitcoin/signature.c, bitcoin/pubkey.c,
 * bitcoin/shadouble.c, common/node_id.c, gossipd/sigcheck.c and wire/fromwire.c are compiled where they
 * lie under /root/reference (recipe: SURVEY.md §8(c)3+4, oracle/Makefile target `cln`) on top of the
 * libwally+libsecp256k1 amalgamation.  This file supplies (a) the handful of externs that normally come
 * from common/utils.c (which needs libsodium) — the same trick the reference's unit tests use with their
 * auto-generated stubs — and (b) flat C entry points for ctypes.  Output: oracle/_ref/libcln_ref.so.
 */
#include "config.h"
#include <bitcoin/pubkey.h>
#include <bitcoin/shadouble.h>
#include <bitcoin/signature.h>
#include <ccan/tal/str/str.h>
#include <common/amount.h>
#include <common/node_id.h>
#include <gossipd/sigcheck.h>
#include <secp256k1.h>
#include <stdlib.h>
#include <string.h>
#include <wire/wire.h>

/* ---- what common/utils.c / common/setup.c would provide ---- */
secp256k1_context *secp256k1_ctx;
const tal_t *tmpctx;
const tal_t *wally_tal_ctx;
const struct chainparams *chainparams;
bool is_elements(const struct chainparams *cp) { (void)cp; return false; }
void tal_wally_start(void) {}
void tal_wally_end(const tal_t *parent) { (void)parent; }
/* the harness's transactions carry no PSBT: the amount of the input being signed is set by the test (cln_tx_set_input_amount) */
static u64 g_input_amount_sat;
struct amount_sat psbt_input_get_amount(const struct wally_psbt *psbt, size_t in) {
    (void)psbt; (void)in;
    struct amount_sat a;
    a.satoshis = g_input_amount_sat;
    return a;
}
bool utf8_check(const void *buf, size_t len) { (void)buf; (void)len; return true; }
char *tal_hexstr(const tal_t *ctx, const void *data, size_t len) {
    static const char d[] = "0123456789abcdef";
    char *s = tal_arr(ctx, char, len * 2 + 1);
    for (size_t i = 0; i < len; i++) { s[2 * i] = d[((const u8 *)data)[i] >> 4]; s[2 * i + 1] = d[((const u8 *)data)[i] & 15]; }
    s[2 * len] = 0;
    return s;
}
char *tal_hex(const tal_t *ctx, const tal_t *data) { return tal_hexstr(ctx, data, tal_bytelen(data)); }

static void setup(void) {
    if (secp256k1_ctx) return;
    secp256k1_ctx = secp256k1_context_create(SECP256K1_CONTEXT_VERIFY | SECP256K1_CONTEXT_SIGN); /* common/setup.c:58 */
    tmpctx = tal(NULL, char);
}
static void sweep(void) { tal_free(tmpctx); tmpctx = tal(NULL, char); }

/* parse exactly as CLN's wire layer does (wire/fromwire.c:188-199, bitcoin/pubkey.c:14,102), then call the
 * reference entry point.  Returns 1/0 = verdict, -1 = the wire layer refuses the encoding. */
int cln_check_signed_hash(const u8 *hash32, const u8 *sig64, const u8 *pub33) {
    setup();
    secp256k1_ecdsa_signature sig;
    struct pubkey key;
    struct sha256_double h;
    const u8 *p = sig64;
    size_t max = 64;
    fromwire_secp256k1_ecdsa_signature(&p, &max, &sig);
    if (!p) return -1;
    if (!pubkey_from_der(pub33, 33, &key)) return -1;
    memcpy(h.sha.u.u8, hash32, 32);
    return check_signed_hash(&h, &sig, &key);
}
int cln_check_signed_hash_nodeid(const u8 *hash32, const u8 *sig64, const u8 *node_id33) {
    setup();
    secp256k1_ecdsa_signature sig;
    struct node_id id;
    struct sha256_double h;
    const u8 *p = sig64;
    size_t max = 64;
    fromwire_secp256k1_ecdsa_signature(&p, &max, &sig);
    if (!p) return -1;
    memcpy(id.k, node_id33, 33);
    memcpy(h.sha.u.u8, hash32, 32);
    return check_signed_hash_nodeid(&h, &sig, &id);
}
int cln_check_schnorr_sig(const u8 *hash32, const u8 *pub33, const u8 *sig64) {
    setup();
    struct pubkey key;
    struct sha256 h;
    struct bip340sig s;
    if (!pubkey_from_der(pub33, 33, &key)) return -1;
    memcpy(h.u.u8, hash32, 32);
    memcpy(s.u8, sig64, 64);
    return check_schnorr_sig(&h, &key.pubkey, &s);
}
void cln_sha256_double(const u8 *p, size_t len, u8 *out32) {
    struct sha256_double h;
    sha256_double(&h, p, len);
    memcpy(out32, h.sha.u.u8, 32);
}
/* opaque structs as CLN's parsers build them (for driving the engine's drop-in entry points) */
int cln_make_opaque(const u8 *sig64, const u8 *pub33, u8 *sig_opaque64, u8 *pub_opaque64) {
    setup();
    secp256k1_ecdsa_signature sig;
    struct pubkey key;
    const u8 *p = sig64;
    size_t max = 64;
    fromwire_secp256k1_ecdsa_signature(&p, &max, &sig);
    if (!p || !pubkey_from_der(pub33, 33, &key)) return 0;
    memcpy(sig_opaque64, sig.data, 64);
    memcpy(pub_opaque64, key.pubkey.data, 64);
    return 1;
}

static int which(const char *err) {
    if (!err) return 0;
    if (strstr(err, "node_signature_1")) return 1;
    if (strstr(err, "node_signature_2")) return 2;
    if (strstr(err, "bitcoin_signature_1")) return 3;
    if (strstr(err, "bitcoin_signature_2")) return 4;
    return 1;
}
/* gossipd/sigcheck.c:45-115 on a raw channel_announcement: 0 ok, 1..4 first bad signature, -1 malformed.
 * Field offsets per wire/peer_wire.csv:340-352 (the generated fromwire_channel_announcement is not in the tree). */
int cln_sigcheck_channel_announcement(const u8 *msg, size_t len) {
    setup();
    if (len < 260) return -1;
    size_t flen = ((size_t)msg[258] << 8) | msg[259], keys = 260 + flen + 32 + 8;
    if (len < keys + 4 * 33) return -1;
    secp256k1_ecdsa_signature sig[4];
    for (int k = 0; k < 4; k++) {
        const u8 *p = msg + 2 + 64 * k;
        size_t max = 64;
        fromwire_secp256k1_ecdsa_signature(&p, &max, &sig[k]);
        if (!p) return -1;
    }
    struct node_id id1, id2;
    struct pubkey b1, b2;
    memcpy(id1.k, msg + keys, 33);
    memcpy(id2.k, msg + keys + 33, 33);
    if (!pubkey_from_der(msg + keys + 66, 33, &b1) || !pubkey_from_der(msg + keys + 99, 33, &b2)) return -1;
    u8 *ann = tal_dup_arr(tmpctx, u8, msg, len, 0);
    const char *err = sigcheck_channel_announcement(tmpctx, &id1, &id2, &b1, &b2, &sig[0], &sig[1], &sig[2], &sig[3], ann);
    int r = which(err);
    sweep();
    return r;
}
int cln_sigcheck_node_announcement(const u8 *msg, size_t len) {
    setup();
    if (len < 68) return -1;
    size_t flen = ((size_t)msg[66] << 8) | msg[67], idoff = 68 + flen + 4;
    /* the generated fromwire_node_announcement (wire/peer_wire.csv:353-362) also pulls rgb_color(3), alias(32),
     * addrlen(2) and addrlen bytes of addresses: a shorter message fails to parse */
    if (len < idoff + 33 + 3 + 32 + 2) return -1;
    size_t alen = ((size_t)msg[idoff + 68] << 8) | msg[idoff + 69];
    if (len < idoff + 70 + alen) return -1;
    secp256k1_ecdsa_signature sig;
    const u8 *p = msg + 2;
    size_t max = 64;
    fromwire_secp256k1_ecdsa_signature(&p, &max, &sig);
    if (!p) return -1;
    struct node_id id;
    memcpy(id.k, msg + idoff, 33);
    u8 *ann = tal_dup_arr(tmpctx, u8, msg, len, 0);
    const char *err = sigcheck_node_announcement(tmpctx, &id, &sig, ann);
    int r = err ? 1 : 0;
    sweep();
    return r;
}
int cln_sigcheck_channel_update(const u8 *msg, size_t len, const u8 *node_id33) {
    setup();
    if (len < 138) return -1; /* fixed layout of wire/peer_wire.csv:366-377 incl. the mandatory htlc_maximum_msat */
    secp256k1_ecdsa_signature sig;
    const u8 *p = msg + 2;
    size_t max = 64;
    fromwire_secp256k1_ecdsa_signature(&p, &max, &sig);
    if (!p) return -1;
    struct node_id id;
    memcpy(id.k, node_id33, 33);
    u8 *upd = tal_dup_arr(tmpctx, u8, msg, len, 0);
    const char *err = sigcheck_channel_update(tmpctx, &id, &sig, upd);
    int r = err ? 1 : 0;
    sweep();
    return r;
}

/* ---- BIP143 sighash exactly as bitcoin_tx_hash_for_sig obtains it (bitcoin/signature.c:120-151): build the
 * one-input one-output transaction with libwally and call wally_tx_get_btc_signature_hash(..., USE_WITNESS).
 * Returns 0 on success (WALLY_OK). ---- */
#include <wally_transaction.h>
int cln_htlc_sighash(uint32_t version, uint32_t locktime, const u8 *prev_txid32, uint32_t prev_index, uint32_t sequence,
                     const u8 *wscript, size_t wscript_len, uint64_t input_amount, uint64_t output_amount,
                     const u8 *out_script, size_t out_script_len, uint32_t sighash_type, u8 *out32) {
    struct wally_tx *tx = NULL;
    int rc = wally_tx_init_alloc(version, locktime, 1, 1, &tx);
    if (rc) return rc;
    rc = wally_tx_add_raw_input(tx, prev_txid32, 32, prev_index, sequence, NULL, 0, NULL, 0);
    if (!rc) rc = wally_tx_add_raw_output(tx, output_amount, out_script, out_script_len, 0);
    if (!rc) rc = wally_tx_get_btc_signature_hash(tx, 0, wscript, wscript_len, input_amount, sighash_type,
                                                  WALLY_TX_FLAG_USE_WITNESS, out32, 32);
    wally_tx_free(tx);
    return rc;
}

/* ---- check_tx_sig itself (bitcoin/signature.c:194-221), on a struct bitcoin_tx assembled with libwally: any number of
 * inputs and outputs, so that commitment transactions (BOLT #3 Appendix C, channeld/test/run-commit_tx.c) fit ---- */
#include <bitcoin/tx.h>
struct bitcoin_tx *cln_tx_new(uint32_t version, uint32_t locktime) {
    setup();
    struct bitcoin_tx *tx = calloc(1, sizeof(*tx));
    if (wally_tx_init_alloc(version, locktime, 4, 4, &tx->wtx) != WALLY_OK) abort();
    return tx;
}
int cln_tx_add_input(struct bitcoin_tx *tx, const u8 *txid32, uint32_t index, uint32_t sequence) {
    return wally_tx_add_raw_input(tx->wtx, txid32, 32, index, sequence, NULL, 0, NULL, 0);
}
int cln_tx_add_output(struct bitcoin_tx *tx, uint64_t satoshi, const u8 *script, size_t len) {
    return wally_tx_add_raw_output(tx->wtx, satoshi, script, len, 0);
}
void cln_tx_free(struct bitcoin_tx *tx) { wally_tx_free(tx->wtx); free(tx); }
void cln_tx_set_input_amount(uint64_t sat) { g_input_amount_sat = sat; }
uint64_t cln_tx_input_amount_hook(const struct bitcoin_tx *tx, size_t in) { (void)tx; (void)in; return g_input_amount_sat; }
size_t cln_tal_bytelen_hook(const void *p) { return tal_bytelen(p); }
u8 *cln_tal_bytes(const u8 *p, size_t len) { return tal_dup_arr(NULL, u8, p, len, 0); }
void cln_tal_free(void *p) { tal_free(p); }
size_t cln_sizeof_bitcoin_signature(void) { return sizeof(struct bitcoin_signature); }
/* fills a struct bitcoin_signature and a struct pubkey the way CLN's parsers would; 0 if either fails to parse */
int cln_make_tx_sig_args(const u8 *sig64, uint32_t sighash_type, const u8 *pub33, void *bitcoin_signature_out, void *pubkey_out) {
    setup();
    struct bitcoin_signature *bs = bitcoin_signature_out;
    struct pubkey *pk = pubkey_out;
    const u8 *p = sig64;
    size_t max = 64;
    fromwire_secp256k1_ecdsa_signature(&p, &max, &bs->s);
    if (!p) return 0;
    bs->sighash_type = (enum sighash_type)sighash_type;
    return pubkey_from_der(pub33, 33, pk) ? 1 : 0;
}
void cln_tx_sighash(const struct bitcoin_tx *tx, unsigned in, const u8 *tal_script, uint32_t sighash_type, u8 *out32) {
    struct sha256_double h;
    bitcoin_tx_hash_for_sig(tx, in, tal_script, (enum sighash_type)sighash_type, &h);
    memcpy(out32, h.sha.u.u8, 32);
}
/* CLN's own, unmodified check_tx_sig */
int cln_check_tx_sig(const struct bitcoin_tx *tx, size_t in, const u8 *tal_redeemscript, const u8 *tal_witness_script,
                     const void *pubkey, const void *bitcoin_signature) {
    return check_tx_sig(tx, in, tal_redeemscript, tal_witness_script, pubkey, bitcoin_signature) ? 1 : 0;
}
