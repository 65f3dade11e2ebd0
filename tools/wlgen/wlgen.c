/* wlgen.c -- synthetic workloads of BASELINE.json's configs 3, 4 and 5.
 *
 * Bench and test infrastructure, not product code.  Integer-only, so that a C++ / Rust caller
 * reproduces every byte; every stream draws from its own SplitMix64 generator seeded from
 * (seed, stream index) alone, so a batch of n streams is a prefix of any longer batch of the same
 * seed ("prefix-stable") and streams can be generated in any order and in parallel.
 *
 *   splitmix64:  s += 0x9E3779B97F4A7C15; z = s;
 *                z = (z ^ z >> 30) * 0xBF58476D1CE4E5B9; z = (z ^ z >> 27) * 0x94D049BB133111EB;
 *                return z ^ z >> 31
 *   stream state: s_i = mix(seed + 0xD1B54A32D192ED03 * (i + 1)), mix = one splitmix64 output
 *
 * Stream i of a workload always consumes its generator in the order written below.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define WL_API __attribute__((visibility("default")))

typedef struct { uint64_t s; } sm64;

static inline uint64_t sm64_next(sm64* g) {
    uint64_t z = (g->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline sm64 stream_gen(uint64_t seed, uint64_t i) {
    sm64 g = { seed + 0xD1B54A32D192ED03ull * (i + 1) };
    sm64 h = { sm64_next(&g) };
    return h;
}

/* a byte source on top of the 64-bit generator: 8 bytes per draw, least significant first */
typedef struct { sm64 g; uint64_t w; int left; } bytesrc;
static inline uint32_t next_byte(bytesrc* b) {
    if (b->left == 0) { b->w = sm64_next(&b->g); b->left = 8; }
    const uint32_t v = (uint32_t)(b->w & 0xFF);
    b->w >>= 8; b->left--;
    return v;
}

/* ---- a small parallel-for over streams (pthreads; OpenMP is not usable with every gcc of the image) */
typedef void (*wl_item_fn)(void* ctx, uint64_t k);
typedef struct { wl_item_fn fn; void* ctx; uint64_t n, grain; uint64_t next; } wl_job;
static void* wl_worker(void* p) {
    wl_job* j = (wl_job*)p;
    for (;;) {
        const uint64_t a = __atomic_fetch_add(&j->next, j->grain, __ATOMIC_RELAXED);
        if (a >= j->n) break;
        const uint64_t b = a + j->grain < j->n ? a + j->grain : j->n;
        for (uint64_t k = a; k < b; k++) j->fn(j->ctx, k);
    }
    return NULL;
}
static void wl_parallel_for(uint64_t n, uint64_t grain, wl_item_fn fn, void* ctx) {
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((uint64_t)nt > (n + grain - 1) / grain) nt = (long)((n + grain - 1) / grain);
    wl_job j = { fn, ctx, n, grain, 0 };
    pthread_t th[64];
    int started = 0;
    for (long t = 1; t < nt; t++)
        if (pthread_create(&th[started], NULL, wl_worker, &j) == 0) started++;
    wl_worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ---- config 3: TIFF strips ----------------------------------------------------------------------
 * len_i = lo + first draw % span; class i % 4:
 *   0 uniform random bytes
 *   1 photo: three interleaved channels, each a random walk, start = one byte, step = byte % 7 - 3
 *   2 runs: every byte starts a new run with probability 1/16 (byte & 15 == 0), value = next byte
 *   3 16-symbol Zipf(1.2) text over " etaoinshrdlucmf": a 16-bit draw against a cumulative table
 */
static const uint16_t kZipfCdf[16] = { 23940, 34360, 40766, 45301, 48771, 51560, 53877, 55851,
                                       57565, 59076, 60423, 61637, 62739, 63748, 64677, 65535 };
static const char kZipfSym[17] = " etaoinshrdlucmf";

WL_API uint64_t wl_tiff_strip_len(uint64_t seed, uint64_t i, uint64_t lo, uint64_t span) {
    sm64 g = stream_gen(seed, i);
    return lo + sm64_next(&g) % span;
}

static void tiff_strip_fill(uint64_t seed, uint64_t i, uint64_t lo, uint64_t span, uint8_t* dst) {
    sm64 g = stream_gen(seed, i);
    const uint64_t n = lo + sm64_next(&g) % span;
    bytesrc b = { g, 0, 0 };
    switch (i & 3) {
    case 0:
        for (uint64_t k = 0; k < n; k++) dst[k] = (uint8_t)next_byte(&b);
        break;
    case 1: {
        uint32_t ch[3];
        for (int c = 0; c < 3; c++) ch[c] = next_byte(&b);
        for (uint64_t k = 0; k < n; k++) {
            const int c = (int)(k % 3);
            ch[c] = (ch[c] + next_byte(&b) % 7 + 256 - 3) & 0xFF;
            dst[k] = (uint8_t)ch[c];
        }
        break;
    }
    case 2: {
        uint32_t v = next_byte(&b);
        for (uint64_t k = 0; k < n; k++) {
            if ((next_byte(&b) & 15) == 0) v = next_byte(&b);
            dst[k] = (uint8_t)v;
        }
        break;
    }
    default:
        for (uint64_t k = 0; k < n; k++) {
            const uint32_t u = next_byte(&b) | (next_byte(&b) << 8);
            int s = 0;
            while (u > kZipfCdf[s]) s++;
            dst[k] = (uint8_t)kZipfSym[s];
        }
        break;
    }
}

/* off[n + 1] must hold the exclusive prefix sum of wl_tiff_strip_len; streams [first, first + n) */
typedef struct { uint64_t seed, first, a, b; const uint64_t* off; uint8_t* buf; const uint8_t* blob;
                 const uint32_t* woff; uint32_t nwords; uint64_t* lens; } wl_args;
static void tiff_item(void* c, uint64_t k) {
    const wl_args* a = (const wl_args*)c;
    tiff_strip_fill(a->seed, a->first + k, a->a, a->b, a->buf + a->off[k]);
}
WL_API void wl_tiff_strips_fill(uint64_t seed, uint64_t first, uint64_t n, uint64_t lo, uint64_t span,
                                const uint64_t* off, uint8_t* buf) {
    wl_args a = { seed, first, lo, span, off, buf, NULL, NULL, 0, NULL };
    wl_parallel_for(n, 64, tiff_item, &a);
}

/* ---- config 4: GIF frames -----------------------------------------------------------------------
 * frame i: cs = 2 + i % 7, pixels < 2^cs, side * side of them.  i % 8 == 7: uniform noise.
 * Otherwise runs (a new run starts with probability 1/8: byte & 7 == 0, value = next byte mod
 * colours) with dither: a second byte < 26 (about 10 %) replaces the pixel by a random colour.
 */
static void gif_frame_fill(uint64_t seed, uint64_t i, uint64_t px, uint8_t* dst) {
    sm64 g = stream_gen(seed, i);
    bytesrc b = { g, 0, 0 };
    const uint32_t cm = (1u << (2 + i % 7)) - 1;
    if (i % 8 == 7) {
        for (uint64_t k = 0; k < px; k++) dst[k] = (uint8_t)(next_byte(&b) & cm);
        return;
    }
    uint32_t v = next_byte(&b) & cm;
    for (uint64_t k = 0; k < px; k++) {
        if ((next_byte(&b) & 7) == 0) v = next_byte(&b) & cm;
        uint32_t p = v;
        if (next_byte(&b) < 26) p = next_byte(&b) & cm;
        dst[k] = (uint8_t)p;
    }
}

static void gif_item(void* c, uint64_t k) {
    const wl_args* a = (const wl_args*)c;
    gif_frame_fill(a->seed, a->first + k, a->a, a->buf + k * a->a);
}
WL_API void wl_gif_frames_fill(uint64_t seed, uint64_t first, uint64_t n, uint64_t px, uint8_t* buf) {
    wl_args a = { seed, first, px, 0, NULL, buf, NULL, NULL, 0, NULL };
    wl_parallel_for(n, 1, gif_item, &a);
}

/* ---- config 5: lorem-like text chunks -------------------------------------------------------------
 * words: `nwords` words, word w = blob[woff[w] .. woff[w + 1]).  Chunk i: words drawn uniformly
 * (draw % nwords), separated by one space; after at least 12 words since the last full stop a
 * draw's bit 0 decides whether ". " follows and the next word is capitalised; a newline replaces
 * the space once the line has 80 characters or more.  The chunk is cut at `chunk` bytes.
 */
static void text_chunk_fill(uint64_t seed, uint64_t i, uint64_t chunk, const uint8_t* blob,
                            const uint32_t* woff, uint32_t nwords, uint8_t* dst) {
    sm64 g = stream_gen(seed, i);
    uint64_t n = 0;
    uint32_t line = 0, since = 0;
    int cap = 1;
    while (n < chunk) {
        const uint32_t w = (uint32_t)(sm64_next(&g) % nwords);
        const uint8_t* p = blob + woff[w];
        const uint32_t l = woff[w + 1] - woff[w];
        for (uint32_t k = 0; k < l && n < chunk; k++) {
            uint8_t c = p[k];
            if (k == 0 && cap && c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
            dst[n++] = c;
        }
        cap = 0;
        line += l;
        since++;
        if (since >= 12 && (sm64_next(&g) & 1)) {
            if (n < chunk) dst[n++] = '.';
            line++;
            since = 0;
            cap = 1;
        }
        if (n < chunk) {
            if (line >= 80) { dst[n++] = '\n'; line = 0; }
            else { dst[n++] = ' '; line++; }
        }
    }
}

static void text_item(void* c, uint64_t k) {
    const wl_args* a = (const wl_args*)c;
    text_chunk_fill(a->seed, a->first + k, a->a, a->blob, a->woff, a->nwords, a->buf + k * a->a);
}
WL_API void wl_text_chunks_fill(uint64_t seed, uint64_t first, uint64_t n, uint64_t chunk, const uint8_t* blob,
                                const uint32_t* woff, uint32_t nwords, uint8_t* buf) {
    wl_args a = { seed, first, chunk, 0, NULL, buf, blob, woff, nwords, NULL };
    wl_parallel_for(n, 16, text_item, &a);
}

WL_API void wl_tiff_strip_lens(uint64_t seed, uint64_t first, uint64_t n, uint64_t lo, uint64_t span, uint64_t* lens) {
    for (uint64_t k = 0; k < n; k++) lens[k] = wl_tiff_strip_len(seed, first + k, lo, span);
}
