/* TEST-ONLY: exposes the generated C codec of the verifier subdaemon's wire messages to ctypes (tests/test_wire_codec.py) */
#include "../../lightning_b200/csrc/sigverifyd_wiregen.h"

size_t shim_towire_verify(uint8_t *out, size_t cap, uint64_t req_id, uint8_t kind, uint32_t n, const uint8_t *hashes,
                          uint32_t keylen, const uint8_t *keys, const uint8_t *sigs) {
    return towire_sigverifyd_verify(out, cap, req_id, kind, n, hashes, keylen, keys, sigs);
}
/* returns 1 and fills the scalar fields + offsets of the views, 0 if the message does not parse */
int shim_fromwire_verify(const uint8_t *p, size_t len, uint64_t *req_id, uint32_t *kind_n_keylen, size_t *offs) {
    struct sigverifyd_verify v;
    if (!fromwire_sigverifyd_verify(p, len, &v)) return 0;
    *req_id = v.req_id;
    kind_n_keylen[0] = v.kind; kind_n_keylen[1] = v.n; kind_n_keylen[2] = v.keylen;
    offs[0] = (size_t)(v.hashes - p); offs[1] = (size_t)(v.keys - p); offs[2] = (size_t)(v.sigs - p);
    return 1;
}
size_t shim_towire_verify_reply(uint8_t *out, size_t cap, uint64_t req_id, uint32_t n, const uint8_t *verdicts) {
    return towire_sigverifyd_verify_reply(out, cap, req_id, n, verdicts);
}
size_t shim_towire_stats_reply(uint8_t *out, size_t cap, uint64_t req_id, uint64_t a, uint64_t b, uint64_t c, uint32_t d) {
    return towire_sigverifyd_stats_reply(out, cap, req_id, a, b, c, d);
}
int shim_fromwire_gossip(const uint8_t *p, size_t len, uint64_t *req_id, uint32_t *n_bloblen, size_t *offs) {
    struct sigverifyd_gossip g;
    if (!fromwire_sigverifyd_gossip(p, len, &g)) return 0;
    *req_id = g.req_id;
    n_bloblen[0] = g.n; n_bloblen[1] = g.bloblen;
    offs[0] = (size_t)(g.lens - p); offs[1] = (size_t)(g.signers - p); offs[2] = (size_t)(g.blob - p);
    return 1;
}
