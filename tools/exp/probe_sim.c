/* probe_sim.c -- CPU simulation of per-lane ("thread per stream") dictionary layouts for the
 * encoder: rounds (dependent table loads) per input byte on config-3 strips, TIFF flavour.
 * Schemes: linear probing over B-slot buckets, double hashing over B-slot buckets, two-choice
 * over B-slot buckets.  One round = the loads a lane can issue at once (one bucket, or both
 * buckets of a pair).  Build: gcc -O2 -o probe_sim probe_sim.c ../wlgen/wlgen.c -pthread */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
uint64_t wl_tiff_strip_len(uint64_t, uint64_t, uint64_t, uint64_t);
void wl_tiff_strips_fill(uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, const uint64_t*, uint8_t*);
#define SLOTS 4096
static uint32_t scr(uint32_t c) { return (c * 0x9E5u) & 0xFFF; }
static uint32_t tbl[SLOTS];
typedef struct { int B, mode; } scheme; /* mode 0 linear, 1 double, 2 two-choice */
static uint64_t rounds, lookups, inserts, maxr;
static uint64_t hist[64];
/* returns code or 0xFFFFFFFF after inserting (key -> code) */
static uint32_t lookup_insert(scheme s, uint32_t prefix, uint32_t byte, uint32_t next_code, int can_insert) {
    const int nb = SLOTS / s.B;
    const uint32_t key = (scr(prefix) << 20) | (byte << 12);
    const uint32_t hb = (byte * 0x6A7u);
    uint32_t h1 = (scr(prefix) ^ hb) % nb;
    uint32_t h2 = ((scr(prefix) * 0x3D5u >> 3) ^ (hb >> 2) ^ (byte * 0x2F1u >> 1)) % nb;
    uint32_t step = 1;
    if (s.mode == 1) step = (((byte * 0x35u) ^ (scr(prefix) >> 5)) % nb) | 1;
    if (s.mode == 3) step = (((byte * 0x9Bu) >> 2) % nb) | 1; /* byte-only step */
    if (s.mode == 4) step = ((scr(prefix) >> 3) % nb) | 1; /* prefix-only step */
    uint32_t r = 0;
    lookups++;
    for (;;) {
        r++;
        uint32_t* b1 = tbl + (size_t)h1 * s.B;
        uint32_t* b2 = tbl + (size_t)h2 * s.B;
        int c1 = 0, c2 = 0;
        for (int i = 0; i < s.B; i++) {
            if (b1[i] && (b1[i] >> 12) == (key >> 12)) { rounds += r; hist[r < 63 ? r : 63]++; if (r > maxr) maxr = r; return b1[i] & 0xFFF; }
            if (b1[i]) c1++;
        }
        if (s.mode == 2) {
            for (int i = 0; i < s.B; i++) {
                if (b2[i] && (b2[i] >> 12) == (key >> 12)) { rounds += r; hist[r < 63 ? r : 63]++; if (r > maxr) maxr = r; return b2[i] & 0xFFF; }
                if (b2[i]) c2++;
            }
        }
        int space = c1 < s.B || (s.mode == 2 && c2 < s.B);
        if (space) {
            rounds += r; hist[r < 63 ? r : 63]++; if (r > maxr) maxr = r;
            if (can_insert) {
                inserts++;
                if (s.mode == 2 && (c2 < c1)) b2[c2] = key | next_code; else if (c1 < s.B) b1[c1] = key | next_code; else b2[c2] = key | next_code;
            }
            return 0xFFFFFFFFu;
        }
        h1 = (h1 + step) % nb;
        h2 = (h2 + step) % nb;
    }
}
int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 0) : 256;
    const uint64_t seed = 0x5A172E60 + 3;
    uint64_t* off = calloc(n + 1, 8);
    for (uint64_t i = 0; i < n; i++) off[i + 1] = off[i] + wl_tiff_strip_len(seed, i, 8192, 57345);
    uint8_t* buf = malloc(off[n]);
    wl_tiff_strips_fill(seed, 0, n, 8192, 57345, off, buf);
    scheme S[] = { {8,1},{8,3},{8,4},{4,2},{4,3},{16,3} };
    for (unsigned si = 0; si < sizeof S / sizeof S[0]; si++) {
        for (int cls = -1; cls < 0; cls++) {
            rounds = lookups = inserts = maxr = 0; memset(hist, 0, sizeof hist);
            uint64_t bytes = 0, codes = 0;
            for (uint64_t i = 0; i < n; i++) {
                if (cls >= 0 && (int)(i & 3) != cls) continue;
                const uint8_t* p = buf + off[i]; const uint64_t len = off[i + 1] - off[i];
                memset(tbl, 0, sizeof tbl);
                uint32_t next = 258, prefix = p[0]; codes++;
                for (uint64_t k = 1; k < len; k++) {
                    uint32_t c = lookup_insert(S[si], prefix, p[k], scr(next), 1);
                    if (c != 0xFFFFFFFFu) { /* stored scrambled */
                        /* unscramble */ prefix = (c * 0xBEDu) & 0xFFF;
                    } else {
                        codes++; prefix = p[k];
                        /* TIFF: entry index next; when next == 4094 after insert -> clear */
                        if (next == 4094) { memset(tbl, 0, sizeof tbl); next = 258; codes++; } else next++;
                    }
                    bytes++;
                }
            }
            printf("B=%2d mode=%d class=%2d  rounds/byte=%.3f  inserts/byte=%.3f max=%llu  r1=%.3f r2=%.3f r3=%.3f r4+=%.3f\n", S[si].B, S[si].mode, cls,
                   (double)rounds / bytes, (double)inserts / bytes, (unsigned long long)maxr,
                   (double)hist[1] / lookups, (double)hist[2] / lookups, (double)hist[3] / lookups,
                   1.0 - (double)(hist[1] + hist[2] + hist[3]) / lookups);
        }
    }
    return 0;
}
