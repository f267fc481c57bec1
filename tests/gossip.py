"""Test-side gossip helpers: parse the committed gossip fixture into signature items exactly the way
gossipd/sigcheck.c (reference) slices the wire messages."""
import json
import os
import struct

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WIRE_CHANNEL_ANNOUNCEMENT, WIRE_NODE_ANNOUNCEMENT, WIRE_CHANNEL_UPDATE = 256, 257, 258


def load_subset():
    blob = open(os.path.join(GOLD, "gossip_subset.bin"), "rb").read()
    msgs, pos = [], 0
    while pos < len(blob):
        (ln,) = struct.unpack(">H", blob[pos:pos + 2])
        msgs.append(blob[pos + 2:pos + 2 + ln])
        pos += 2 + ln
    return msgs


def items_of(msgs):
    """-> (data uint8 array, off uint64[], len uint32[], key33[n,33], sig[n,64], owner[n] (message index),
    which[n] (0..3 = position of the signature inside its message))"""
    chans = {}
    data = np.frombuffer(b"".join(msgs), dtype=np.uint8)
    starts = np.cumsum([0] + [len(m) for m in msgs])[:-1]
    off, ln, keys, sigs, owner, which = [], [], [], [], [], []
    for mi, m in enumerate(msgs):
        typ = struct.unpack(">H", m[:2])[0]
        if typ == WIRE_CHANNEL_ANNOUNCEMENT:  # sigcheck.c:45-115: hash msg[258:], 4 sigs at 2+64k
            flen = struct.unpack(">H", m[258:260])[0]
            p = 260 + flen + 32
            scid = m[p:p + 8]
            p += 8
            ks = [m[p + 33 * k:p + 33 * k + 33] for k in range(4)]
            chans[scid] = (ks[0], ks[1])
            for k in range(4):
                off.append(starts[mi] + 258); ln.append(len(m) - 258); keys.append(ks[k])
                sigs.append(m[2 + 64 * k:66 + 64 * k]); owner.append(mi); which.append(k)
        elif typ == WIRE_NODE_ANNOUNCEMENT:  # sigcheck.c:118-164: hash msg[66:]
            flen = struct.unpack(">H", m[66:68])[0]
            p = 68 + flen + 4
            off.append(starts[mi] + 66); ln.append(len(m) - 66); keys.append(m[p:p + 33])
            sigs.append(m[2:66]); owner.append(mi); which.append(0)
        elif typ == WIRE_CHANNEL_UPDATE:  # sigcheck.c:9-43: hash msg[66:], signer chosen by channel_flags & 1
            scid = m[98:106]
            chflags = m[111]
            if scid in chans:
                off.append(starts[mi] + 66); ln.append(len(m) - 66); keys.append(chans[scid][chflags & 1])
                sigs.append(m[2:66]); owner.append(mi); which.append(0)
    n = len(off)
    key = np.frombuffer(b"".join(keys), dtype=np.uint8).reshape(n, 33).copy()
    sig = np.frombuffer(b"".join(sigs), dtype=np.uint8).reshape(n, 64).copy()
    return (data.copy(), np.array(off, np.uint64), np.array(ln, np.uint32), key, sig, np.array(owner), np.array(which))


def chan_ann_3703():
    j = json.load(open(os.path.join(GOLD, "chan_ann_3703.json")))
    return bytes.fromhex(j["msg"])


def strip_features(m):
    """Re-encode a channel_announcement with an empty feature field (what the reference test does)."""
    flen = struct.unpack(">H", m[258:260])[0]
    return m[:258] + b"\x00\x00" + m[260 + flen:]
