/*
 * slzw_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY, NOT PART OF THE PRODUCT.
 *
 * A literal C restatement of the hot path of redwarp/lzw (crate salzweg 0.1.3), keeping the
 * reference's own data structures (3-state arena trie; prefix/suffix/length tables + word
 * stack; byte-granular bit I/O through a 32-bit accumulator) so that it doubles as the timed
 * "salzweg CPU path" stand-in.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file; lzw_b200/ never does.
 *
 * The reference is Rust and there is no Rust toolchain in this image (no cargo/rustc), so
 * oracle/_ref cannot be built; parity of this restatement is pinned by the reference's own
 * known-answer vectors instead (tests/test_oracle_golden.py): lorem_ipsum.txt <->
 * lorem_ipsum_encoded.bin (encoder.rs:740-755, decoder.rs:703-718), the 40-byte 4-colour
 * vectors (encoder.rs:666-686, 798-813), [0,0,1,3] for GIF/TIFF/fixed (encoder.rs:689-712,
 * 816-835), the bit-I/O vectors (io.rs:334-572) and the error vectors (encoder.rs:758-795,
 * decoder.rs:721-737, 759-769).
 *
 * Each function cites the reference lines it follows.  All citations are relative to
 * /root/reference/lzw/src/.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/slzw.h"

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * std::io stand-ins.  `src_t` is `&[u8] as Read`; `sink_t` is `&mut [u8] as Write` (or a Vec
 * when cap is the encode bound; or a byte counter when p == NULL).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* p;
    size_t n;
    size_t pos;
} src_t;

typedef struct {
    uint8_t* p; /* NULL = count only (unbounded Vec whose bytes nobody reads) */
    size_t cap;
    size_t len;
} sink_t;

/* Write::write_all on `&mut [u8]`: copies what fits, then fails with WriteZero. */
static inline int sink_write_all(sink_t* s, const uint8_t* b, size_t n) {
    if (!s->p) {
        s->len += n;
        return 0;
    }
    size_t room = s->cap - s->len;
    size_t m = n < room ? n : room;
    memcpy(s->p + s->len, b, m);
    s->len += m;
    return m < n ? SLZW_ERR_IO_WRITE_ZERO : 0;
}

/* ------------------------------------------------------------------------------------------
 * io.rs:205-328  LittleEndianWriter / BigEndianWriter
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    sink_t* w;
    uint8_t cursor;
    uint32_t byte_buffer;
} bitw_t;

/* io.rs:234-248 */
static inline int le_write(bitw_t* b, uint16_t data, uint8_t amount) {
    uint32_t mask = (1u << amount) - 1;
    b->byte_buffer |= ((uint32_t)data & mask) << b->cursor;
    b->cursor += amount;
    while (b->cursor >= 8) {
        uint8_t byte = (uint8_t)b->byte_buffer;
        b->byte_buffer >>= 8;
        b->cursor -= 8;
        int e = sink_write_all(b->w, &byte, 1);
        if (e) return e;
    }
    return 0;
}

/* io.rs:251-259 */
static inline int le_fill(bitw_t* b) {
    if (b->cursor > 0) {
        uint8_t byte = (uint8_t)b->byte_buffer;
        int e = sink_write_all(b->w, &byte, 1);
        if (e) return e;
        b->byte_buffer = 0;
        b->cursor = 0;
    }
    return 0;
}

/* io.rs:296-311 */
static inline int be_write(bitw_t* b, uint16_t data, uint8_t amount) {
    uint32_t mask = (1u << amount) - 1;
    uint32_t shift = 32u - amount - b->cursor;
    b->byte_buffer |= ((uint32_t)data & mask) << shift;
    b->cursor += amount;
    while (b->cursor >= 8) {
        uint8_t byte = (uint8_t)(b->byte_buffer >> 24);
        b->byte_buffer <<= 8;
        b->cursor -= 8;
        int e = sink_write_all(b->w, &byte, 1);
        if (e) return e;
    }
    return 0;
}

/* io.rs:314-322 */
static inline int be_fill(bitw_t* b) {
    if (b->cursor > 0) {
        uint8_t byte = (uint8_t)(b->byte_buffer >> 24);
        int e = sink_write_all(b->w, &byte, 1);
        if (e) return e;
        b->byte_buffer = 0;
        b->cursor = 0;
    }
    return 0;
}

static inline int bw_write(bitw_t* b, int big, uint16_t data, uint8_t amount) {
    return big ? be_write(b, data, amount) : le_write(b, data, amount);
}
static inline int bw_fill(bitw_t* b, int big) { return big ? be_fill(b) : le_fill(b); }

/* ------------------------------------------------------------------------------------------
 * io.rs:11-153  LittleEndianReader / BigEndianReader
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    src_t* r;
    uint8_t cursor;
    uint32_t byte_buffer;
} bitr_t;

/* io.rs:43-55 (read_one: read_exact => EOF is an error) and io.rs:58-78 (read: EOF => short
 * count).  Returns 1 on success, 0 on EOF; the caller decides what EOF means. */
static inline int le_read_one(bitr_t* b, uint8_t amount, uint16_t* out) {
    while (b->cursor < amount) {
        if (b->r->pos >= b->r->n) return 0;
        uint8_t byte = b->r->p[b->r->pos++];
        b->byte_buffer |= (uint32_t)byte << b->cursor;
        b->cursor += 8;
    }
    uint32_t mask = (1u << amount) - 1;
    *out = (uint16_t)(b->byte_buffer & mask);
    b->byte_buffer >>= amount;
    b->cursor -= amount;
    return 1;
}

/* io.rs:113-128 and io.rs:131-152 */
static inline int be_read_one(bitr_t* b, uint8_t amount, uint16_t* out) {
    while (b->cursor < amount) {
        if (b->r->pos >= b->r->n) return 0;
        uint8_t byte = b->r->p[b->r->pos++];
        uint32_t shift = 24u - b->cursor;
        b->byte_buffer |= (uint32_t)byte << shift;
        b->cursor += 8;
    }
    uint32_t mask = (1u << amount) - 1;
    uint32_t shift = 32u - amount;
    *out = (uint16_t)((b->byte_buffer >> shift) & mask);
    b->byte_buffer <<= amount;
    b->cursor -= amount;
    return 1;
}

static inline int br_read_one(bitr_t* b, int big, uint8_t amount, uint16_t* out) {
    return big ? be_read_one(b, amount, out) : le_read_one(b, amount, out);
}

/* ------------------------------------------------------------------------------------------
 * encoder.rs:58-149  Node / Tree
 * ---------------------------------------------------------------------------------------- */
enum { NODE_NO_CHILD = 0, NODE_ONE_CHILD = 1, NODE_MANY_CHILDREN = 2 };

typedef struct {
    uint8_t tag;
    uint8_t child_char;   /* OneChild.0 */
    uint16_t child_index; /* OneChild.1 */
    uint16_t* children;   /* ManyChildren(Vec<u16>) */
} node_t;

#define MAX_ENTRY_COUNT 4097 /* encoder.rs:76 */

typedef struct {
    node_t nodes[MAX_ENTRY_COUNT];
    size_t len;
    uint8_t code_size;
    size_t code_count;
    int with_clear_code;
} tree_t;

/* encoder.rs:75-85 */
static void tree_new(tree_t* t, uint8_t code_size, int with_clear_code) {
    t->len = 0;
    t->code_size = code_size;
    t->code_count = (size_t)1 << code_size;
    t->with_clear_code = with_clear_code;
}

/* encoder.rs:88-95  (nodes.clear() drops every ManyChildren Vec) */
static void tree_reset(tree_t* t) {
    for (size_t i = 0; i < t->len; i++)
        if (t->nodes[i].tag == NODE_MANY_CHILDREN) free(t->nodes[i].children);
    size_t n = ((size_t)1 << t->code_size) + (t->with_clear_code ? 2 : 0);
    for (size_t i = 0; i < n; i++) t->nodes[i].tag = NODE_NO_CHILD;
    t->len = n;
}

static void tree_drop(tree_t* t) {
    for (size_t i = 0; i < t->len; i++)
        if (t->nodes[i].tag == NODE_MANY_CHILDREN) free(t->nodes[i].children);
    t->len = 0;
}

/* encoder.rs:98-118.  Returns 1 = Some(*word), 0 = None, -1 = the reference panics
 * (prefix_index out of bounds, only reachable through the unchecked first byte). */
static inline int tree_find_word(const tree_t* t, uint16_t prefix_index, uint8_t next_char,
                                 uint16_t* word) {
    if (prefix_index >= t->len) return -1;
    const node_t* prefix = &t->nodes[prefix_index];
    switch (prefix->tag) {
        case NODE_NO_CHILD:
            return 0;
        case NODE_ONE_CHILD:
            if (prefix->child_char == next_char) {
                *word = prefix->child_index;
                return 1;
            }
            return 0;
        default: {
            uint16_t child_index = prefix->children[next_char];
            if (child_index > 0) {
                *word = child_index;
                return 1;
            }
            return 0;
        }
    }
}

/* encoder.rs:121-143 */
static inline uint16_t tree_add(tree_t* t, uint16_t prefix_index, uint8_t k) {
    uint16_t new_index = (uint16_t)t->len;
    node_t* old_node = &t->nodes[prefix_index];
    switch (old_node->tag) {
        case NODE_NO_CHILD:
            old_node->tag = NODE_ONE_CHILD;
            old_node->child_char = k;
            old_node->child_index = new_index;
            break;
        case NODE_ONE_CHILD: {
            uint16_t* children = (uint16_t*)calloc(t->code_count, sizeof(uint16_t));
            children[old_node->child_char] = old_node->child_index;
            children[k] = new_index;
            old_node->tag = NODE_MANY_CHILDREN;
            old_node->children = children;
            break;
        }
        default:
            old_node->children[k] = new_index;
            break;
    }
    t->nodes[t->len].tag = NODE_NO_CHILD;
    t->len++;
    return new_index;
}

/* ------------------------------------------------------------------------------------------
 * encoder.rs:273-346  VariableEncoder::inner_encode
 * ---------------------------------------------------------------------------------------- */
static int variable_encode(src_t* data, sink_t* into, uint8_t code_size, int big, int tiff,
                           uint32_t* detail, tree_t* tree) {
    const uint8_t MAX_WRITE_SIZE = 12;
    *detail = 0;
    if (!(code_size >= 2 && code_size <= 8)) { /* 281-283 */
        *detail = code_size;
        return SLZW_ERR_CODE_SIZE;
    }
    uint8_t max_code = (uint8_t)((1u << code_size) - 1); /* 285 */
    bitw_t bw = {into, 0, 0};
    uint16_t increment = tiff ? 1 : 0; /* lib.rs:84-91 */
    int e;

    uint8_t write_size = code_size + 1;                                  /* 289 */
    uint16_t clear_code = (uint16_t)(1u << code_size);                   /* 290 */
    uint16_t end_of_information = (uint16_t)((1u << code_size) + 1);     /* 291 */
    uint16_t size_increase_mask = (uint16_t)((1u << write_size) - increment); /* 292 */

    tree_new(tree, code_size, 1); /* 294-295 */
    tree_reset(tree);

    if ((e = bw_write(&bw, big, clear_code, write_size))) return e; /* 297 */

    if (data->pos >= data->n) { /* 299-309: empty stream */
        if ((e = bw_write(&bw, big, end_of_information, write_size))) return e;
        if ((e = bw_fill(&bw, big))) return e;
        return SLZW_OK;
    }

    uint16_t current_prefix = data->p[data->pos++]; /* 311: first byte is NOT range-checked */

    while (data->pos < data->n) { /* 313 */
        uint8_t k = data->p[data->pos++];
        if (k > max_code) { /* 315-317 */
            *detail = k;
            return SLZW_ERR_UNEXPECTED_CODE;
        }
        uint16_t word;
        int found = tree_find_word(tree, current_prefix, k, &word); /* 319 */
        if (found < 0) return SLZW_ERR_REFERENCE_PANIC;          /* encoder.rs:99 index panic */
        if (found) {
            current_prefix = word; /* 320 */
        } else {
            uint16_t index_of_new_entry = tree_add(tree, current_prefix, k);    /* 322 */
            if ((e = bw_write(&bw, big, current_prefix, write_size))) return e; /* 323 */
            current_prefix = k;                                                 /* 324 */

            if (index_of_new_entry == size_increase_mask) { /* 326 */
                if (write_size < MAX_WRITE_SIZE) {
                    write_size += 1;
                } else {
                    if ((e = bw_write(&bw, big, clear_code, MAX_WRITE_SIZE))) return e; /* 330 */
                    write_size = code_size + 1;
                    tree_reset(tree);
                }
                size_increase_mask = (uint16_t)((1u << write_size) - increment); /* 334 */
            }
        }
    }

    if ((e = bw_write(&bw, big, current_prefix, write_size))) return e;     /* 339 */
    if ((e = bw_write(&bw, big, end_of_information, write_size))) return e; /* 340 */
    if ((e = bw_fill(&bw, big))) return e;                                  /* 342 */
    return SLZW_OK;
}

/* ------------------------------------------------------------------------------------------
 * encoder.rs:618-658  FixedEncoder::inner_encode
 * ---------------------------------------------------------------------------------------- */
static int fixed_encode(src_t* data, sink_t* into, int big, uint32_t* detail, tree_t* tree) {
    const uint8_t WRITE_SIZE = 12;
    const size_t MAX_TABLE_SIZE = 4096;
    *detail = 0;
    bitw_t bw = {into, 0, 0};
    int e;

    tree_new(tree, 8, 0); /* 624-625 */
    tree_reset(tree);

    if (data->pos >= data->n) { /* 629-635 */
        if ((e = bw_fill(&bw, big))) return e;
        return SLZW_OK;
    }

    uint16_t current_prefix = data->p[data->pos++]; /* 637 */

    while (data->pos < data->n) { /* 639 */
        uint8_t k = data->p[data->pos++];
        uint16_t word;
        int found = tree_find_word(tree, current_prefix, k, &word); /* 642 */
        if (found < 0) return SLZW_ERR_REFERENCE_PANIC;          /* unreachable: 256 roots */
        if (found) {
            current_prefix = word;
        } else {
            if (tree->len < MAX_TABLE_SIZE) tree_add(tree, current_prefix, k);  /* 645-647 */
            if ((e = bw_write(&bw, big, current_prefix, WRITE_SIZE))) return e; /* 648 */
            current_prefix = k;
        }
    }

    if ((e = bw_write(&bw, big, current_prefix, WRITE_SIZE))) return e; /* 653 */
    if ((e = bw_fill(&bw, big))) return e;                              /* 654 */
    return SLZW_OK;
}

/* ------------------------------------------------------------------------------------------
 * decoder.rs:174-290  VariableDecoder::inner_decode
 * ---------------------------------------------------------------------------------------- */
#define MAX_TABLE_SIZE 4096 /* decoder.rs:185 */
#define MAX_STACK_SIZE 4091 /* decoder.rs:192 */

typedef struct {
    uint16_t prefix[MAX_TABLE_SIZE];
    uint8_t suffix[MAX_TABLE_SIZE];
    size_t length[MAX_TABLE_SIZE];
    uint8_t decoding_stack[MAX_STACK_SIZE];
} dec_tables_t;

/* `lenient` is NOT in the reference: it restates include/slzw.h's SLZW_FLAVOUR_VARIABLE_LENIENT
 * (a full dictionary freezes instead of failing) so that the extension has a checker too. */
static int variable_decode(src_t* data, sink_t* into, uint8_t code_size, int big, int tiff,
                           int lenient, uint32_t* detail, dec_tables_t* t) {
    const uint8_t MAX_READ_SIZE = 12;
    *detail = 0;
    if (!(code_size >= 2 && code_size <= 8)) { /* 180-182 */
        *detail = code_size;
        return SLZW_ERR_CODE_SIZE;
    }
    uint16_t increment = tiff ? 1 : 0;
    int e;

    memset(t->prefix, 0, sizeof t->prefix); /* 197-201 */
    memset(t->suffix, 0, sizeof t->suffix);
    memset(t->length, 0, sizeof t->length);
    memset(t->decoding_stack, 0, sizeof t->decoding_stack);
    for (uint32_t code = 0; code < (1u << code_size); code++) { /* 203-206 */
        t->suffix[code] = (uint8_t)code;
        t->length[code] = 1;
    }

    uint8_t read_size = code_size + 1;                                       /* 208 */
    uint16_t clear_code = (uint16_t)(1u << code_size);                       /* 210 */
    uint16_t end_of_information = clear_code + 1;                            /* 211 */
    uint16_t size_increase_mask = (uint16_t)((1u << read_size) - increment); /* 213 */
    uint16_t next_index = clear_code + 2;                                    /* 214 */
    int have_previous = 0;                                                   /* 215: Option */
    uint16_t previous_code = 0;
    bitr_t br = {data, 0, 0};
    size_t word_length = 0; /* 217 */

    for (;;) { /* 219 */
        uint16_t code;
        if (!br_read_one(&br, big, read_size, &code)) return SLZW_ERR_IO_UNEXPECTED_EOF; /* 220 */

        if (code == clear_code) { /* 222-227: tables are NOT cleared */
            read_size = code_size + 1;
            size_increase_mask = (uint16_t)((1u << read_size) - increment);
            next_index = clear_code + 2;
            have_previous = 0;
            continue;
        } else if (code == end_of_information) { /* 228-229 */
            break;
        } else if (!have_previous) { /* 230-236: no range check on `code` */
            if ((e = sink_write_all(into, &t->suffix[code], 1))) return e;
            have_previous = 1;
            previous_code = code;
            t->decoding_stack[0] = (uint8_t)code;
            word_length = 1;
            continue;
        }

        uint16_t initial_code = code; /* 238 */

        if (code > next_index) { /* 241-243 */
            *detail = code;
            return SLZW_ERR_UNEXPECTED_CODE;
        } else if (code == next_index) { /* 244-250 */
            if (word_length >= MAX_STACK_SIZE) return SLZW_ERR_REFERENCE_PANIC; /* index panic */
            t->decoding_stack[word_length] = t->decoding_stack[0];
            word_length += 1;
        } else { /* 251-267 */
            word_length = t->length[code];
            size_t stack_top = word_length;
            while (code >= clear_code) { /* 256 */
                stack_top -= 1;
                if (stack_top == 0) { /* 258-260 */
                    *detail = code;
                    return SLZW_ERR_UNEXPECTED_CODE;
                }
                if (stack_top >= MAX_STACK_SIZE) return SLZW_ERR_REFERENCE_PANIC; /* index panic */
                t->decoding_stack[stack_top] = t->suffix[code]; /* 262 */
                code = t->prefix[code];                         /* 263 */
            }
            t->decoding_stack[0] = (uint8_t)code; /* 266 */
        }

        if (word_length > MAX_STACK_SIZE) return SLZW_ERR_REFERENCE_PANIC; /* slice panic, 270 */
        if ((e = sink_write_all(into, t->decoding_stack, word_length))) return e; /* 270 */

        if (next_index < MAX_TABLE_SIZE) { /* 272-280 */
            t->prefix[next_index] = previous_code;
            t->suffix[next_index] = t->decoding_stack[0];
            t->length[next_index] = t->length[previous_code] + 1;
            next_index += 1;
            if (next_index == size_increase_mask && read_size < MAX_READ_SIZE) {
                read_size += 1;
                size_increase_mask = (uint16_t)((1u << read_size) - increment);
            }
        } else if (!lenient) {
            return SLZW_ERR_MISSING_CLEAR_CODE; /* 281-283 */
        }
        previous_code = initial_code; /* 284 */
    }
    return SLZW_OK; /* 287-289 */
}

/* ------------------------------------------------------------------------------------------
 * decoder.rs:553-642  FixedDecoder::inner_decode
 * ---------------------------------------------------------------------------------------- */
static int fixed_decode(src_t* data, sink_t* into, int big, uint32_t* detail, dec_tables_t* t) {
    const uint8_t READ_SIZE = 12;
    *detail = 0;
    int e;

    memset(t->prefix, 0, sizeof t->prefix); /* 569-573 */
    memset(t->suffix, 0, sizeof t->suffix);
    memset(t->length, 0, sizeof t->length);
    memset(t->decoding_stack, 0, sizeof t->decoding_stack);
    for (uint32_t code = 0; code < 256; code++) { /* 575-578 */
        t->suffix[code] = (uint8_t)code;
        t->length[code] = 1;
    }

    uint16_t next_index = 256; /* 580 */
    int have_previous = 0;
    uint16_t previous_code = 0;
    bitr_t br = {data, 0, 0};
    size_t word_length = 0;

    for (;;) { /* 585: bit_reader.iter(12) ends on a short read (io.rs:58-64, 183-194) */
        uint16_t code;
        if (!br_read_one(&br, big, READ_SIZE, &code)) break;

        if (!have_previous) { /* 588-594 */
            if ((e = sink_write_all(into, &t->suffix[code], 1))) return e;
            have_previous = 1;
            previous_code = code;
            t->decoding_stack[0] = (uint8_t)code;
            word_length = 1;
            continue;
        }

        uint16_t initial_code = code; /* 596 */

        if (code > next_index) { /* 599-601 */
            *detail = code;
            return SLZW_ERR_UNEXPECTED_CODE;
        } else if (code == next_index) { /* 602-608 */
            if (word_length >= MAX_STACK_SIZE) return SLZW_ERR_REFERENCE_PANIC;
            t->decoding_stack[word_length] = t->decoding_stack[0];
            word_length += 1;
        } else { /* 609-625 */
            word_length = t->length[code];
            size_t stack_top = word_length;
            while (code >= 256) { /* 614 */
                stack_top -= 1;
                if (stack_top == 0) { /* 616-618 */
                    *detail = code;
                    return SLZW_ERR_UNEXPECTED_CODE;
                }
                if (stack_top >= MAX_STACK_SIZE) return SLZW_ERR_REFERENCE_PANIC;
                t->decoding_stack[stack_top] = t->suffix[code];
                code = t->prefix[code];
            }
            t->decoding_stack[0] = (uint8_t)code; /* 624 */
        }

        if (word_length > MAX_STACK_SIZE) return SLZW_ERR_REFERENCE_PANIC;
        if ((e = sink_write_all(into, t->decoding_stack, word_length))) return e; /* 628 */

        if (next_index < MAX_TABLE_SIZE) { /* 630-635: full table simply stops growing */
            t->prefix[next_index] = previous_code;
            t->suffix[next_index] = t->decoding_stack[0];
            t->length[next_index] = t->length[previous_code] + 1;
            next_index += 1;
        }
        previous_code = initial_code; /* 636 */
    }
    return SLZW_OK;
}

/* ------------------------------------------------------------------------------------------
 * Facade dispatch: encoder.rs:199-220, 392-399, 479-487, 565-576;
 *                  decoder.rs:99-120, 333-340, 420-428, 503-514.
 * out == NULL counts bytes only (cap ignored).  Returns slzw_status.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_encode(const slzw_params* params, const uint8_t* in, uint64_t n,
                             uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail) {
    tree_t* tree = (tree_t*)malloc(sizeof(tree_t));
    src_t src = {in, (size_t)n, 0};
    sink_t sink = {out, (size_t)cap, 0};
    uint32_t d = 0;
    int st;
    tree->len = 0;
    if (params->flavour == SLZW_FLAVOUR_FIXED)
        st = fixed_encode(&src, &sink, params->big_endian != 0, &d, tree);
    else
        st = variable_encode(&src, &sink, params->code_size, params->big_endian != 0,
                             params->tiff_early_change != 0, &d, tree);
    tree_drop(tree);
    free(tree);
    if (out_len) *out_len = sink.len;
    if (detail) *detail = d;
    return st;
}

ORACLE_API int oracle_decode(const slzw_params* params, const uint8_t* in, uint64_t n,
                             uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail) {
    dec_tables_t* t = (dec_tables_t*)malloc(sizeof(dec_tables_t));
    src_t src = {in, (size_t)n, 0};
    sink_t sink = {out, (size_t)cap, 0};
    uint32_t d = 0;
    int st;
    if (params->flavour == SLZW_FLAVOUR_FIXED)
        st = fixed_decode(&src, &sink, params->big_endian != 0, &d, t);
    else
        st = variable_decode(&src, &sink, params->code_size, params->big_endian != 0,
                             params->tiff_early_change != 0,
                             params->flavour == SLZW_FLAVOUR_VARIABLE_LENIENT, &d, t);
    free(t);
    if (out_len) *out_len = sink.len;
    if (detail) *detail = d;
    return st;
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers for the CPU baseline: one stream per task across `threads` host threads (the
 * stand-in for rayon par_iter over streams).  Same slzw_batch convention as
 * include/slzw.h with host pointers.  Each worker reuses one tree / one table set, like one
 * rayon worker calling the stateless reference functions back to back.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const slzw_params* params;
    const slzw_batch* b;
    int decode;
    atomic_ullong* next;
} job_t;

static void* batch_worker(void* arg) {
    job_t* j = (job_t*)arg;
    const slzw_batch* b = j->b;
    tree_t* tree = NULL;
    dec_tables_t* tabs = NULL;
    if (j->decode)
        tabs = (dec_tables_t*)malloc(sizeof(dec_tables_t));
    else {
        tree = (tree_t*)malloc(sizeof(tree_t));
        tree->len = 0;
    }
    for (;;) {
        /* grab 16 streams at a time */
        unsigned long long i0 = atomic_fetch_add(j->next, 16ull);
        if (i0 >= b->n) break;
        unsigned long long i1 = i0 + 16 < b->n ? i0 + 16 : b->n;
        for (unsigned long long i = i0; i < i1; i++) {
            slzw_params p = *j->params;
            if (b->code_size) p.code_size = b->code_size[i];
            src_t src = {b->in + b->in_off[i], (size_t)(b->in_off[i + 1] - b->in_off[i]), 0};
            sink_t sink = {b->out ? b->out + b->out_off[i] : NULL,
                           b->out ? (size_t)(b->out_off[i + 1] - b->out_off[i]) : 0, 0};
            uint32_t d = 0;
            int st;
            if (j->decode) {
                if (p.flavour == SLZW_FLAVOUR_FIXED)
                    st = fixed_decode(&src, &sink, p.big_endian != 0, &d, tabs);
                else
                    st = variable_decode(&src, &sink, p.code_size, p.big_endian != 0,
                                         p.tiff_early_change != 0,
                                         p.flavour == SLZW_FLAVOUR_VARIABLE_LENIENT, &d, tabs);
            } else {
                if (p.flavour == SLZW_FLAVOUR_FIXED)
                    st = fixed_encode(&src, &sink, p.big_endian != 0, &d, tree);
                else
                    st = variable_encode(&src, &sink, p.code_size, p.big_endian != 0,
                                         p.tiff_early_change != 0, &d, tree);
                tree_drop(tree);
            }
            if (b->out_len) b->out_len[i] = sink.len;
            if (b->status) b->status[i] = (uint32_t)st;
            if (b->detail) b->detail[i] = d;
        }
    }
    free(tree);
    free(tabs);
    return NULL;
}

static int run_batch(const slzw_params* params, const slzw_batch* b, int decode, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    atomic_ullong next;
    atomic_init(&next, 0);
    job_t job = {params, b, decode, &next};
    if (threads == 1) {
        batch_worker(&job);
        return 0;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int i = 0; i < threads; i++) {
        if (pthread_create(&th[i], NULL, batch_worker, &job) != 0) break;
        started++;
    }
    if (started == 0) batch_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

ORACLE_API int oracle_encode_batch(const slzw_params* params, const slzw_batch* batch,
                                   int threads) {
    return run_batch(params, batch, 0, threads);
}

ORACLE_API int oracle_decode_batch(const slzw_params* params, const slzw_batch* batch,
                                   int threads) {
    return run_batch(params, batch, 1, threads);
}

/* ------------------------------------------------------------------------------------------
 * Bit I/O exposed for the io.rs known-answer vectors (io.rs:334-572).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API uint64_t oracle_bitwrite(int big, const uint16_t* codes, const uint8_t* widths,
                                    uint64_t n, uint8_t* out, uint64_t cap) {
    sink_t sink = {out, (size_t)cap, 0};
    bitw_t bw = {&sink, 0, 0};
    for (uint64_t i = 0; i < n; i++)
        if (bw_write(&bw, big, codes[i], widths[i])) return sink.len;
    bw_fill(&bw, big);
    return sink.len;
}

/* reads codes of the given widths with read_one; returns how many were read before EOF */
ORACLE_API uint64_t oracle_bitread(int big, const uint8_t* in, uint64_t n_in,
                                   const uint8_t* widths, uint64_t n, uint16_t* codes) {
    src_t src = {in, (size_t)n_in, 0};
    bitr_t br = {&src, 0, 0};
    uint64_t i = 0;
    for (; i < n; i++)
        if (!br_read_one(&br, big, widths[i], &codes[i])) break;
    return i;
}
