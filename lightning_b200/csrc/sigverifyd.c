/*
 * sigverifyd.c — verifier subdaemon: ONE GPU-owning process that serves signature
 * verification to the other daemons over a unix socket, in the shape of CLN's own subdaemons
 * (lightningd/subd.c:796 new_global_subd; single-threaded poll loop like ccan/io; length-prefixed binary
 * frames like the generated wire messages).  It replaces "one CUDA context per daemon" (SURVEY.md §8b).
 *
 *   cln_sigverifyd <socket-path> [cuda-device]
 *
 * Frame (all integers big-endian, as on CLN's wires):
 *   request : u32 len | u8 kind | u32 n | msg32[n] | key[n * keysize(kind)] | sig64[n]
 *   reply   : u32 len | u32 n | verdict[n]            (verdict 0/1; an engine failure closes the connection)
 * Requests are served in arrival order, one batch launch per request; concurrent clients are multiplexed by poll().
 */
#define _GNU_SOURCE
#include "../../include/cln_sigverify.h"

#include <errno.h>
#include <poll.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#define MAX_CLIENTS 64
#define MAX_FRAME (1u << 30)

typedef struct {
    int fd;
    uint8_t *buf;
    size_t have, want; /* want == 0: reading the 4-byte length */
    uint8_t hdr[4];
} client_t;

static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void put32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

static int write_all(int fd, const uint8_t *p, size_t n) {
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return -1; }
        p += w; n -= (size_t)w;
    }
    return 0;
}

/* returns 0 ok, -1 protocol or engine error (connection is dropped) */
static int serve(sv_ctx *ctx, client_t *c) {
    const uint8_t *p = c->buf;
    size_t len = c->want;
    if (len < 5) return -1;
    int kind = p[0];
    uint32_t n = be32(p + 1);
    size_t ks = sv_key_size(kind);
    if (ks == 0 || len != 5 + (size_t)n * (32 + ks + 64)) return -1;
    const uint8_t *msg = p + 5, *key = msg + 32 * (size_t)n, *sig = key + ks * (size_t)n;
    uint8_t *reply = (uint8_t *)malloc(8 + (size_t)n);
    if (!reply) return -1;
    put32(reply, 4 + n);
    put32(reply + 4, n);
    int rc = sv_verify_host(ctx, kind, msg, key, sig, n, reply + 8);
    if (rc != SV_OK) {
        fprintf(stderr, "cln_sigverifyd: engine error %d: %s\n", rc, sv_last_error(ctx));
        free(reply);
        return -1;
    }
    rc = write_all(c->fd, reply, 8 + (size_t)n);
    free(reply);
    return rc;
}

static void drop(client_t *c) {
    close(c->fd);
    free(c->buf);
    memset(c, 0, sizeof *c);
    c->fd = -1;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <socket-path> [cuda-device]\n", argv[0]); return 2; }
    signal(SIGPIPE, SIG_IGN);
    sv_ctx *ctx = NULL;
    int rc = sv_create(&ctx, argc > 2 ? atoi(argv[2]) : 0);
    if (rc != SV_OK) { fprintf(stderr, "cln_sigverifyd: sv_create failed (%d): %s\n", rc, sv_last_error(NULL)); return 1; }
    int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    struct sockaddr_un addr;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    strncpy(addr.sun_path, argv[1], sizeof addr.sun_path - 1);
    unlink(argv[1]);
    if (ls < 0 || bind(ls, (struct sockaddr *)&addr, sizeof addr) < 0 || listen(ls, 16) < 0) { perror("cln_sigverifyd: socket"); return 1; }
    fprintf(stderr, "cln_sigverifyd: ready on %s\n", argv[1]);
    client_t cl[MAX_CLIENTS];
    for (int i = 0; i < MAX_CLIENTS; i++) { memset(&cl[i], 0, sizeof cl[i]); cl[i].fd = -1; }
    for (;;) {
        struct pollfd pfd[MAX_CLIENTS + 1];
        int idx[MAX_CLIENTS + 1], np = 0;
        pfd[np].fd = ls; pfd[np].events = POLLIN; idx[np++] = -1;
        for (int i = 0; i < MAX_CLIENTS; i++)
            if (cl[i].fd >= 0) { pfd[np].fd = cl[i].fd; pfd[np].events = POLLIN; idx[np++] = i; }
        if (poll(pfd, (nfds_t)np, -1) < 0) { if (errno == EINTR) continue; break; }
        for (int k = 0; k < np; k++) {
            if (!(pfd[k].revents & (POLLIN | POLLHUP | POLLERR))) continue;
            if (idx[k] < 0) {
                int fd = accept(ls, NULL, NULL);
                if (fd < 0) continue;
                int slot = -1;
                for (int i = 0; i < MAX_CLIENTS; i++) if (cl[i].fd < 0) { slot = i; break; }
                if (slot < 0) { close(fd); continue; }
                cl[slot].fd = fd;
                continue;
            }
            client_t *c = &cl[idx[k]];
            if (c->want == 0) { /* length prefix */
                ssize_t r = read(c->fd, c->hdr + c->have, 4 - c->have);
                if (r <= 0) { drop(c); continue; }
                c->have += (size_t)r;
                if (c->have == 4) {
                    c->want = be32(c->hdr);
                    c->have = 0;
                    if (c->want == 0 || c->want > MAX_FRAME || !(c->buf = (uint8_t *)malloc(c->want))) { drop(c); continue; }
                }
            } else {
                ssize_t r = read(c->fd, c->buf + c->have, c->want - c->have);
                if (r <= 0) { drop(c); continue; }
                c->have += (size_t)r;
                if (c->have == c->want) {
                    int s = serve(ctx, c);
                    free(c->buf);
                    c->buf = NULL; c->have = c->want = 0;
                    if (s < 0) drop(c);
                }
            }
        }
    }
    sv_destroy(ctx);
    return 0;
}
