/*
 * cln_dropin.c — host side of the drop-in (plain C, as the reference's bitcoin/signature.c is).
 * It only marshals bytes: opaque libsecp256k1 structs -> wire form, wire messages -> (span, key,
 * signature) items, and calls the batch C ABI.  No arithmetic happens on the host.
 */
#include "../../include/cln_dropin.h"
#include "../../include/cln_sigverify.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if !defined(__BYTE_ORDER__) || __BYTE_ORDER__ != __ORDER_LITTLE_ENDIAN__
#error "opaque-struct conversion below assumes a little-endian 64-bit libsecp256k1 build"
#endif

static sv_ctx *g_ctx;
static int g_device = -1;

static void die(const char *what, int rc) {
    fprintf(stderr, "cln_sigverify: %s failed (%d): %s\n", what, rc, sv_last_error(g_ctx));
    abort(); /* internal error is fatal (CLN convention); never reported as "bad signature" */
}

static sv_ctx *ctx(void) {
    if (!g_ctx) {
        int dev = g_device;
        if (dev < 0) {
            const char *e = getenv("CLN_SIGVERIFY_DEVICE");
            dev = e ? atoi(e) : 0;
        }
        int rc = sv_create(&g_ctx, dev);
        if (rc != SV_OK) {
            fprintf(stderr, "cln_sigverify: sv_create(device %d) failed (%d): %s\n", dev, rc, sv_last_error(NULL));
            abort();
        }
    }
    return g_ctx;
}

void cln_sigverify_init(int device) {
    g_device = device;
    (void)ctx();
}
void cln_sigverify_shutdown(void) {
    if (g_ctx) sv_destroy(g_ctx);
    g_ctx = NULL;
}

/* libsecp256k1's opaque structs hold r,s / x,y as 4x64-bit little-endian limbs on 64-bit little-endian
 * builds (secp256k1.c:337-359, group_impl.h:968-986): 32 bytes little-endian each.  Inside CLN one would
 * call secp256k1_ecdsa_signature_serialize_compact / secp256k1_ec_pubkey_serialize instead
 * (INTEGRATION.md); the engine deliberately does not link libsecp256k1. */
static void rev32(u8 *out, const unsigned char *in) {
    for (int i = 0; i < 32; i++) out[i] = in[31 - i];
}
static void sig_to_wire(u8 out[64], const secp256k1_ecdsa_signature *s) {
    rev32(out, s->data);
    rev32(out + 32, s->data + 32);
}
static void pubkey_to_xy(u8 out[64], const secp256k1_pubkey *p) {
    rev32(out, p->data);
    rev32(out + 32, p->data + 32);
}

bool check_signed_hash(const struct sha256_double *hash, const secp256k1_ecdsa_signature *signature,
                       const struct pubkey *key) {
    u8 sig[64], xy[64], v = 0;
    sig_to_wire(sig, signature);
    pubkey_to_xy(xy, &key->pubkey);
    int rc = sv_verify_host(ctx(), SV_KIND_ECDSA_XY, hash->sha.u.u8, xy, sig, 1, &v);
    if (rc != SV_OK) die("sv_verify_host", rc);
    return v == 1;
}

bool check_signed_hash_nodeid(const struct sha256_double *hash, const secp256k1_ecdsa_signature *signature,
                              const struct node_id *id) {
    u8 sig[64], v = 0;
    sig_to_wire(sig, signature);
    int rc = sv_verify_host(ctx(), SV_KIND_ECDSA33, hash->sha.u.u8, id->k, sig, 1, &v);
    if (rc != SV_OK) die("sv_verify_host", rc);
    return v == 1;
}

bool check_schnorr_sig(const struct sha256 *hash, const secp256k1_pubkey *pubkey, const struct bip340sig *sig) {
    /* signature.c:412-423: serialize compressed, drop the parity byte -> x-only key */
    u8 xy[64], v = 0;
    pubkey_to_xy(xy, pubkey);
    int rc = sv_verify_host(ctx(), SV_KIND_SCHNORR, hash->u.u8, xy, sig->u8, 1, &v);
    if (rc != SV_OK) die("sv_verify_host", rc);
    return v == 1;
}

void sha256_double(struct sha256_double *shadouble, const void *p, size_t len) {
    uint64_t off = 0;
    uint32_t l = (uint32_t)len;
    u8 dummy = 0;
    int rc = sv_sha256d_host(ctx(), len ? (const u8 *)p : &dummy, len, &off, &l, 1, shadouble->sha.u.u8);
    if (rc != SV_OK) die("sv_sha256d_host", rc);
}

bool pubkey_from_der(const u8 *der, size_t len, struct pubkey *key) {
    if (len != 33) return false; /* PUBKEY_CMPR_LEN, bitcoin/pubkey.c:16 */
    u8 xy[64], ok = 0;
    int rc = sv_pubkey_parse_host(ctx(), der, 1, xy, &ok);
    if (rc != SV_OK) die("sv_pubkey_parse_host", rc);
    if (!ok) return false;
    rev32(key->pubkey.data, xy);
    rev32(key->pubkey.data + 32, xy + 32);
    return true;
}

void check_tx_sigs_batch(const struct sha256_double *hashes, const struct bitcoin_signature *sigs,
                         const struct pubkey *key, size_t n, bool *ok) {
    if (n == 0) return;
    u8 *buf = (u8 *)malloc(n * (32 + 64 + 64 + 1));
    if (!buf) die("malloc", -3);
    u8 *msg = buf, *xy = buf + 32 * n, *sig = xy + 64 * n, *v = sig + 64 * n;
    for (size_t i = 0; i < n; i++) {
        memcpy(msg + 32 * i, hashes[i].sha.u.u8, 32);
        pubkey_to_xy(xy + 64 * i, &key->pubkey);
        sig_to_wire(sig + 64 * i, &sigs[i].s);
    }
    int rc = sv_verify_host(ctx(), SV_KIND_ECDSA_XY, msg, xy, sig, n, v);
    if (rc != SV_OK) die("sv_verify_host", rc);
    for (size_t i = 0; i < n; i++) ok[i] = v[i] == 1;
    free(buf);
}

/* ---- gossip: slice raw wire messages the way gossipd/sigcheck.c does and verify them as one batch ---- */
typedef struct {
    uint64_t *off;
    uint32_t *len;
    u8 *key, *sig, *verdict;
    size_t *owner;
    size_t n, cap;
} items_t;

static void items_init(items_t *it, size_t cap) {
    it->cap = cap ? cap : 1;
    it->n = 0;
    it->off = (uint64_t *)malloc(it->cap * sizeof(uint64_t));
    it->len = (uint32_t *)malloc(it->cap * sizeof(uint32_t));
    it->key = (u8 *)malloc(it->cap * 33);
    it->sig = (u8 *)malloc(it->cap * 64);
    it->verdict = (u8 *)calloc(it->cap, 1);
    it->owner = (size_t *)malloc(it->cap * sizeof(size_t));
    if (!it->off || !it->len || !it->key || !it->sig || !it->verdict || !it->owner) die("malloc", -3);
}
static void items_free(items_t *it) {
    free(it->off); free(it->len); free(it->key); free(it->sig); free(it->verdict); free(it->owner);
}
static void items_add(items_t *it, uint64_t off, uint32_t len, const u8 *key33, const u8 *sig64, size_t owner) {
    size_t i = it->n++;
    it->off[i] = off; it->len[i] = len; it->owner[i] = owner;
    memcpy(it->key + 33 * i, key33, 33);
    memcpy(it->sig + 64 * i, sig64, 64);
}
static uint16_t be16(const u8 *p) { return (uint16_t)((p[0] << 8) | p[1]); }

static void run_items(items_t *it, const u8 *blob, size_t blob_len) {
    if (it->n == 0) return;
    int rc = sv_verify_host_raw(ctx(), SV_KIND_ECDSA33, blob, blob_len, it->off, it->len, it->key, it->sig, it->n,
                                it->verdict);
    if (rc != SV_OK) die("sv_verify_host_raw", rc);
}

/* concatenates the messages; returns the blob and fills starts[] */
static u8 *concat(const u8 *const *msgs, const size_t *lens, size_t n, uint64_t *starts, size_t *total) {
    size_t t = 0;
    for (size_t i = 0; i < n; i++) { starts[i] = t; t += lens[i]; }
    u8 *blob = (u8 *)malloc(t ? t : 1);
    if (!blob) die("malloc", -3);
    for (size_t i = 0; i < n; i++) memcpy(blob + starts[i], msgs[i], lens[i]);
    *total = t;
    return blob;
}

void sigcheck_channel_announcement_batch(const u8 *const *msgs, const size_t *lens, size_t n, int *status) {
    /* wire/peer_wire.csv:340-352: type(2) sig x4 (2,66,130,194) flen(2)@258 features chain_hash(32) scid(8)
     * node_id_1 node_id_2 bitcoin_key_1 bitcoin_key_2 (33 each).  Signed region: msg[258:] (sigcheck.c:75). */
    uint64_t *starts = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    size_t total;
    u8 *blob = concat(msgs, lens, n, starts, &total);
    items_t it;
    items_init(&it, 4 * n);
    for (size_t i = 0; i < n; i++) {
        const u8 *m = msgs[i];
        status[i] = 0;
        if (lens[i] < 260 || be16(m) != 256) { status[i] = -1; continue; }
        size_t flen = be16(m + 258), keys = 260 + flen + 32 + 8;
        if (lens[i] < keys + 4 * 33) { status[i] = -1; continue; }
        for (int k = 0; k < 4; k++)
            items_add(&it, starts[i] + 258, (uint32_t)(lens[i] - 258), m + keys + 33 * k, m + 2 + 64 * k, i);
    }
    run_items(&it, blob, total);
    for (size_t j = 0, k = 0; j < it.n; j++) {
        size_t o = it.owner[j];
        k = (j > 0 && it.owner[j - 1] == o) ? k + 1 : 0;
        if (!it.verdict[j] && status[o] == 0) status[o] = (int)k + 1; /* first failure wins (sigcheck.c:79-112) */
    }
    items_free(&it);
    free(blob);
    free(starts);
}

void sigcheck_node_announcement_batch(const u8 *const *msgs, const size_t *lens, size_t n, int *status) {
    /* type(2) sig(64) flen(2)@66 features timestamp(4) node_id(33) ...; signed region msg[66:] (sigcheck.c:141) */
    uint64_t *starts = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    size_t total;
    u8 *blob = concat(msgs, lens, n, starts, &total);
    items_t it;
    items_init(&it, n);
    for (size_t i = 0; i < n; i++) {
        const u8 *m = msgs[i];
        status[i] = 0;
        if (lens[i] < 68 || be16(m) != 257) { status[i] = -1; continue; }
        size_t flen = be16(m + 66), id = 68 + flen + 4;
        if (lens[i] < id + 33) { status[i] = -1; continue; }
        items_add(&it, starts[i] + 66, (uint32_t)(lens[i] - 66), m + id, m + 2, i);
    }
    run_items(&it, blob, total);
    for (size_t j = 0; j < it.n; j++)
        if (!it.verdict[j]) status[it.owner[j]] = 1;
    items_free(&it);
    free(blob);
    free(starts);
}

void sigcheck_channel_update_batch(const u8 *const *msgs, const size_t *lens, const struct node_id *signers,
                                   size_t n, int *status) {
    /* type(2) sig(64) chain_hash(32) scid(8) ...; signed region msg[66:] (sigcheck.c:33) */
    uint64_t *starts = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    size_t total;
    u8 *blob = concat(msgs, lens, n, starts, &total);
    items_t it;
    items_init(&it, n);
    for (size_t i = 0; i < n; i++) {
        status[i] = 0;
        if (lens[i] < 66 + 32 + 8 || be16(msgs[i]) != 258) { status[i] = -1; continue; }
        items_add(&it, starts[i] + 66, (uint32_t)(lens[i] - 66), signers[i].k, msgs[i] + 2, i);
    }
    run_items(&it, blob, total);
    for (size_t j = 0; j < it.n; j++)
        if (!it.verdict[j]) status[it.owner[j]] = 1;
    items_free(&it);
    free(blob);
    free(starts);
}
