/* minigzip_b200.c -- the reference's test/minigzip.c use case on the GPU library: gzip / gunzip a file or stdin to
 * stdout through zlib-ng's own stream API (zng_deflateInit2(level, 31) + zng_deflate, zng_inflateInit2(31) +
 * zng_inflate), linked against libzng_b200.so.  Written for this repository (the reference's tool goes through its
 * gz file layer, gzwrite.c / gzread.c, which is outside the hot path).
 *
 *   minigzip_b200 [-1|-2|-3] [-d] [file]     compressed / decompressed bytes go to stdout
 *
 * Compression feeds the whole input with one zng_deflate(Z_FINISH): the library cuts it into 65536-byte pieces, and the
 * output is byte-identical to the reference fed one piece per zng_deflate(Z_FULL_FLUSH) call (pigz -b 64 -i style).
 */
#include "zlib-ng.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint8_t *slurp(FILE *f, size_t *n) {
    size_t cap = 1 << 20, len = 0;
    uint8_t *p = (uint8_t *)malloc(cap);
    while (p) {
        size_t r = fread(p + len, 1, cap - len, f);
        len += r;
        if (r == 0) break;
        if (len == cap) { cap *= 2; p = (uint8_t *)realloc(p, cap); }
    }
    *n = len;
    return p;
}

int main(int argc, char **argv) {
    int level = 6, decompress = 0;                          /* test/minigzip.c: Z_DEFAULT_COMPRESSION */
    const char *path = NULL;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-d")) decompress = 1;
        else if (argv[i][0] == '-' && argv[i][1] >= '1' && argv[i][1] <= '6' && !argv[i][2]) level = argv[i][1] - '0';
        else if (argv[i][0] == '-' && argv[i][1]) { fprintf(stderr, "usage: %s [-1 .. -6] [-d] [file]\n", argv[0]); return 2; }
        else path = argv[i];
    }
    FILE *f = path && strcmp(path, "-") ? fopen(path, "rb") : stdin;
    if (!f) { perror(path); return 1; }
    size_t n = 0;
    uint8_t *in = slurp(f, &n);
    if (!in) { fprintf(stderr, "out of memory\n"); return 1; }
    zng_stream s;
    memset(&s, 0, sizeof(s));
    if (!decompress) {
        int r = zng_deflateInit2(&s, level, Z_DEFLATED, MAX_WBITS + 16, DEF_MEM_LEVEL, Z_DEFAULT_STRATEGY);
        if (r != Z_OK) { fprintf(stderr, "zng_deflateInit2: %d %s\n", r, s.msg ? s.msg : ""); return 1; }
        size_t cap = zng_deflateBound(&s, (unsigned long)n);
        uint8_t *out = (uint8_t *)malloc(cap);
        size_t off = 0;
        do {                                                  /* avail_in is 32 bits wide: feed <= 1 GiB per call */
            size_t take = n - off > ((size_t)1 << 30) ? ((size_t)1 << 30) : n - off;
            s.next_in = in + off; s.avail_in = (uint32_t)take; off += take;
            do {
                s.next_out = out; s.avail_out = cap > 0xffffffffu ? 0xffffffffu : (uint32_t)cap;
                r = zng_deflate(&s, off == n ? Z_FINISH : Z_FULL_FLUSH);
                if (r != Z_OK && r != Z_STREAM_END) { fprintf(stderr, "zng_deflate: %d %s\n", r, s.msg ? s.msg : ""); return 1; }
                fwrite(out, 1, (size_t)(s.next_out - out), stdout);
            } while (s.avail_out == 0 || s.avail_in != 0);
        } while (r != Z_STREAM_END);
        zng_deflateEnd(&s);
    } else {
        size_t off = 0;
        while (off < n) {                                     /* concatenated members, like gunzip */
            memset(&s, 0, sizeof(s));
            int r = zng_inflateInit2(&s, MAX_WBITS + 16);
            if (r != Z_OK) { fprintf(stderr, "zng_inflateInit2: %d\n", r); return 1; }
            size_t cap = 4 * (n - off) + 65536;
            uint8_t *out = (uint8_t *)malloc(cap);
            s.next_in = in + off; s.avail_in = (uint32_t)(n - off > 0xffffffffu ? 0xffffffffu : n - off);
            do {
                s.next_out = out; s.avail_out = cap > 0xffffffffu ? 0xffffffffu : (uint32_t)cap;
                r = zng_inflate(&s, Z_NO_FLUSH);
                if (r != Z_OK && r != Z_STREAM_END) { fprintf(stderr, "zng_inflate: %d %s\n", r, s.msg ? s.msg : ""); return 1; }
                fwrite(out, 1, (size_t)(s.next_out - out), stdout);
            } while (r != Z_STREAM_END);
            off += s.total_in;
            zng_inflateEnd(&s);
            free(out);
        }
    }
    free(in);
    return fflush(stdout) ? 1 : 0;
}
