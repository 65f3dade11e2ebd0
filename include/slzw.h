/*
 * slzw.h -- C ABI of the B200-native batched LZW codec (salzweg-compatible).
 *
 * This is the drop-in boundary for the hot path of redwarp/lzw (crate `salzweg`):
 * the GIF-style, TIFF-style and fixed-12-bit LZW encoders and decoders.  The
 * reference has no FFI of its own; its boundary is the public Rust API
 *   lzw/src/encoder.rs:199-220, 262-271  VariableEncoder::{encode, encode_to_vec}
 *   lzw/src/encoder.rs:392-399, 435-439  GifStyleEncoder::{encode, encode_to_vec}
 *   lzw/src/encoder.rs:479-487, 519-523  TiffStyleEncoder::{encode, encode_to_vec}
 *   lzw/src/encoder.rs:565-576, 609-616  FixedEncoder::{encode, encode_to_vec}
 *   lzw/src/decoder.rs:99-120, 163-172   VariableDecoder::{decode, decode_to_vec}
 *   lzw/src/decoder.rs:333-340, 378-382  GifStyleDecoder::{decode, decode_to_vec}
 *   lzw/src/decoder.rs:420-428, 460-464  TiffStyleDecoder::{decode, decode_to_vec}
 *   lzw/src/decoder.rs:503-514, 544-551  FixedDecoder::{decode, decode_to_vec}
 * Every entry point below says which of those it backs.  A Rust `salzweg`-shaped
 * crate binds exactly these symbols (see INTEGRATION.md and bindings/rust/).
 *
 * All pointers are plain memory, all sizes are bytes, no C++/torch types cross
 * this boundary.  Functions return SLZW_RC_* (launch-level result); per-stream
 * results are reported through status[]/detail[]/out_len[] and mirror the
 * reference's error enums (lzw/src/encoder.rs:16-29, lzw/src/decoder.rs:15-25).
 *
 * There is no CPU fallback: every entry point that moves data runs CUDA kernels
 * on an sm_100a device and fails with SLZW_RC_NO_DEVICE / SLZW_RC_CUDA otherwise.
 */
#ifndef SLZW_H
#define SLZW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SLZW_API __attribute__((visibility("default")))
#else
#define SLZW_API
#endif

#define SLZW_VERSION_MAJOR 0
#define SLZW_VERSION_MINOR 1

/* ---- per-stream result codes (mirror salzweg's error enums) ------------------------- */
typedef enum slzw_status {
    SLZW_OK = 0,
    /* EncodingError::CodeSize(cs) / DecodingError::CodeSize(cs); detail = cs.
     * (encoder.rs:281-283, decoder.rs:180-182) */
    SLZW_ERR_CODE_SIZE = 1,
    /* EncodingError::UnexpectedCode{code, code_size}: detail = offending input byte
     * (encoder.rs:315-317).  DecodingError::UnexpectedCode(code): detail = code
     * (decoder.rs:241-243, 258-260, 599-601, 616-618). */
    SLZW_ERR_UNEXPECTED_CODE = 2,
    /* DecodingError::MissingClearCode (decoder.rs:281-283). */
    SLZW_ERR_MISSING_CLEAR_CODE = 3,
    /* DecodingError::Io(UnexpectedEof): input ended before the end-of-information
     * code (io.rs:45, 115 via decoder.rs:220). */
    SLZW_ERR_IO_UNEXPECTED_EOF = 4,
    /* {Encoding,Decoding}Error::Io(WriteZero): the output slot is full; it behaves
     * like the `&mut [u8]` writer the reference's benches use
     * (benches/compare_crates.rs:65-77; io.rs:244, 307; decoder.rs:231, 270). */
    SLZW_ERR_IO_WRITE_ZERO = 5,
    /* The reference would panic (index out of bounds) on this input: encoder first
     * byte >= 2^cs+2 followed by more data (encoder.rs:99, 311), or a decoder word
     * longer than its 4091-byte stack (decoder.rs:247, 262).  out_len = bytes the
     * reference had written before panicking. */
    SLZW_ERR_REFERENCE_PANIC = 6
} slzw_status;

/* ---- function-level return codes ----------------------------------------------------- */
typedef enum slzw_rc {
    SLZW_RC_OK = 0,
    SLZW_RC_CUDA = -1,        /* a CUDA call failed; see slzw_last_error() */
    SLZW_RC_INVALID = -2,     /* bad argument (null pointer, bad flavour, ...) */
    SLZW_RC_NO_DEVICE = -3,   /* no usable sm_100 CUDA device */
    SLZW_RC_NOMEM = -4
} slzw_rc;

/* ---- codec options ---------------------------------------------------------------------
 * flavour 0 = variable-width codes with clear/EOI (VariableEncoder/Decoder,
 *             encoder.rs:273-346, decoder.rs:174-290)
 * flavour 1 = fixed 12-bit codes, no control codes (FixedEncoder/Decoder,
 *             encoder.rs:618-658, decoder.rs:553-642)
 * code_size          2..=8, ignored by flavour 1
 * big_endian         0 = Endianness::LittleEndian (LSB-first), 1 = BigEndian (MSB-first)
 * tiff_early_change  0 = CodeSizeStrategy::Default, 1 = CodeSizeStrategy::Tiff
 *                    (lib.rs:71-91), ignored by flavour 1
 * Presets: GIF = {0, cs, 0, 0} (encoder.rs:392-399); TIFF = {0, 8, 1, 1}
 * (encoder.rs:479-487); Fixed = {1, 0, le|be, 0} (encoder.rs:565-576). */
typedef struct slzw_params {
    uint8_t flavour;
    uint8_t code_size;
    uint8_t big_endian;
    uint8_t tiff_early_change;
} slzw_params;

#define SLZW_FLAVOUR_VARIABLE 0
#define SLZW_FLAVOUR_FIXED 1
/* Extension, decoders only: flavour 0, except that a full dictionary is not an
 * error -- the decoder keeps decoding with the 4096 entries it has until a clear code arrives
 * ("deferred clear", which GIF encoders other than salzweg's may use), instead of returning
 * MissingClearCode (decoder.rs:281-283).  Encoders treat it as flavour 0. */
#define SLZW_FLAVOUR_VARIABLE_LENIENT 2

/* ---- a batch of independent streams -----------------------------------------------------
 * Stream i reads in[in_off[i] .. in_off[i+1]) and may write out[out_off[i] .. out_off[i+1])
 * (its capacity slot, the analogue of the reference's `&mut [u8]` writer).  On return
 *   out_len[i] = bytes produced (including those produced before an error),
 *   status[i]  = slzw_status, detail[i] = see slzw_status.
 * code_size (optional, may be NULL) overrides params.code_size per stream (GIF frames of
 * different palettes in one batch).  For the *_device entry points every pointer is a device
 * pointer; for the *_host entry points every pointer is a host pointer. */
typedef struct slzw_batch {
    const uint8_t* in;
    const uint64_t* in_off;   /* n + 1 */
    uint8_t* out;
    const uint64_t* out_off;  /* n + 1 */
    uint64_t* out_len;        /* n */
    uint32_t* status;         /* n */
    uint32_t* detail;         /* n */
    const uint8_t* code_size; /* n or NULL */
    uint64_t n;
} slzw_batch;

typedef struct slzw_ctx slzw_ctx;

/* ---- context ------------------------------------------------------------------------------
 * A context owns the per-device workspace (stream schedule, work queue, staging buffers).
 * One context per host thread; contexts are independent (the reference's functions are
 * stateless and re-entrant).  The *_host entry points of one context run one at a
 * time: a call that arrives while another one is running on the same context returns
 * SLZW_RC_INVALID (its staging buffers are never interleaved); *_device calls of one context are
 * serialised by a lock.  slzw_create fails with SLZW_RC_NO_DEVICE on anything but compute
 * capability 10.0 (the kernels are sm_100a code). */
SLZW_API int slzw_create(int device, slzw_ctx** ctx);
SLZW_API void slzw_destroy(slzw_ctx* ctx);
SLZW_API const char* slzw_last_error(const slzw_ctx* ctx);
/* kernels launched by this context so far (for benchmark accounting) */
SLZW_API uint64_t slzw_kernel_launches(const slzw_ctx* ctx);
SLZW_API uint32_t slzw_version(void);
/* Diagnostics: the streams of the most recent decode call of this context that the fast decode
 * kernel handed to the exact-emulation kernel (malformed streams whose outcome depends on stale
 * table state, more than 1 MiB of output between two clear codes; see decode_kernels.cu).  Synchronises the device.  Returns their number and
 * copies up to `cap` stream ids into `ids` (may be NULL).  After a *_host call that was split
 * into several chunks it describes the last chunk only (ids relative to that chunk). */
SLZW_API uint64_t slzw_last_deferred(slzw_ctx* ctx, uint32_t* ids, uint64_t cap);
/* Diagnostics: input bytes of the most recent encode call of this context by the kind of warp
 * that encoded them -- [0] one warp per stream, dictionary in tensor memory; [1] one warp per
 * stream, dictionary in shared memory; [2] one lane per stream, shared memory; [3] one lane per
 * stream, global memory (encode_kernels.cu).  Synchronises the device.  Device-path and chunked
 * host calls only: SLZW_RC_INVALID after a streaming dense encode (and before any encode call). */
SLZW_API int slzw_last_encode_shares(slzw_ctx* ctx, uint64_t bytes[4]);

/* ---- batched entry points (the hot path; new relative to the reference) ----------------- */
/* Device-resident batch, asynchronous on `cuda_stream` (a cudaStream_t, may be NULL).
 * Each stream is encoded exactly as VariableEncoder::inner_encode (encoder.rs:273-346) or
 * FixedEncoder::inner_encode (encoder.rs:618-658) would encode it on its own. */
SLZW_API int slzw_encode_batch_device(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch,
                             void* cuda_stream);
/* VariableDecoder::inner_decode (decoder.rs:174-290) / FixedDecoder::inner_decode
 * (decoder.rs:553-642) per stream, same conventions. */
SLZW_API int slzw_decode_batch_device(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch,
                             void* cuda_stream);
/* Host-resident batch: copies in, runs the device path, copies results back, synchronous.
 * Buffers from slzw_host_alloc() (pinned) overlap transfers with kernels.
 * The whole capacity range out[out_off[0] .. out_off[n]) is copied back: bytes of a slot beyond
 * out_len[i] are overwritten with unspecified values (the reference's `&mut [u8]` writer leaves
 * them untouched), and an encode call moves the worst-case slots over the bus -- writers that
 * want only the encoded bytes use slzw_encode_batch_host_dense. */
SLZW_API int slzw_encode_batch_host(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch);
SLZW_API int slzw_decode_batch_host(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch);

/* Host-resident batch with DENSE output: stream i's encoded bytes land at
 * out_dense[out_off[i] .. out_off[i+1]) (out_off has n+1 entries and is written by the call,
 * out_off[0] = 0, every stream rounded up to `align` bytes).  Worst-case slots live on the
 * device only; the compaction stage runs before the copy back, so only encoded bytes cross the
 * bus.  If out_off[n] would exceed out_cap the contents of out_dense are unspecified (streams
 * that fit in front of the overflow may have been copied), *needed (if not NULL) receives the
 * required size and SLZW_RC_NOMEM is returned.  The call keeps input + worst-case slots + dense
 * output on the device (4.2 x the input bytes, context-owned, grow-only) and runs as one
 * streaming launch fed by the copy engine; calls above 12 GiB of input are chunked.  This is what a TIFF/GIF writer wants:
 * strips back to back plus StripOffsets/StripByteCounts. */
SLZW_API int slzw_encode_batch_host_dense(slzw_ctx* ctx, const slzw_params* params,
                                          const uint8_t* in, const uint64_t* in_off, uint64_t n,
                                          const uint8_t* code_size, uint64_t align,
                                          uint8_t* out_dense, uint64_t out_cap, uint64_t* out_off,
                                          uint32_t* status, uint32_t* detail, uint64_t* needed);

/* The same in two phases, for callers that have to know the encoded size before they can say
 * where the bytes go (a writer that allocates exactly; several devices filling one dense buffer,
 * slzw_multi_*).  _begin encodes and compacts; the encoded bytes stay on the device, out_off[n+1]
 * (relative to the first stream, out_off[0] = 0), status and detail come back and *total receives
 * out_off[n].  _finish copies the `total` bytes to out_dense (SLZW_RC_NOMEM if out_cap is smaller;
 * the bytes then stay available for another _finish).  A new _begin, or any other dense encode
 * call on the same context, discards what an earlier _begin left behind. */
SLZW_API int slzw_encode_batch_host_dense_begin(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in,
                                       const uint64_t* in_off, uint64_t n, const uint8_t* code_size,
                                       uint64_t align, uint64_t* out_off, uint32_t* status,
                                       uint32_t* detail, uint64_t* total);
SLZW_API int slzw_encode_batch_host_dense_finish(slzw_ctx* ctx, uint8_t* out_dense, uint64_t out_cap);

/* ---- single stream (= batch of one): backs the 16 facade functions ------------------------
 * Returns SLZW_RC_* (<0) on launch failure, otherwise the stream's slzw_status (>=0). */
SLZW_API int slzw_encode(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in, uint64_t n,
                uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail);
SLZW_API int slzw_decode(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in, uint64_t n,
                uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail);

/* ---- sizing ------------------------------------------------------------------------------ */
/* Worst-case encoded size of an n-byte stream (what encode_to_vec needs, encoder.rs:268). */
SLZW_API uint64_t slzw_encode_bound(const slzw_params* params, uint64_t n);
/* Size-only decode: out_len/status/detail as slzw_decode_batch_device with unlimited
 * capacity, nothing written (batch->out/out_off may be NULL).  Backs decode_to_vec
 * (decoder.rs:163-172), which has no caller-supplied capacity. */
SLZW_API int slzw_decoded_sizes_batch_device(slzw_ctx* ctx, const slzw_params* params,
                                    const slzw_batch* batch, void* cuda_stream);
/* The same for a host-resident batch (batch->out/out_off may be NULL). */
SLZW_API int slzw_decoded_sizes_batch_host(slzw_ctx* ctx, const slzw_params* params,
                                  const slzw_batch* batch);

/* ---- compaction (scheduler output stage) -------------------------------------------------
 * dst_off[0..n] = exclusive prefix sum of len[0..n), each length rounded up to a multiple of
 * `align` (1 = dense; TIFF strips want 2); dst[dst_off[i] .. +len[i]) =
 * src[src_off[i] .. +len[i]).  Device pointers, asynchronous on `cuda_stream`. */
SLZW_API int slzw_compact_device(slzw_ctx* ctx, const uint8_t* src, const uint64_t* src_off,
                        const uint64_t* len, uint64_t n, uint64_t align, uint8_t* dst,
                        uint64_t* dst_off, void* cuda_stream);

/* ---- TIFF Predictor = 2 (container step either side of the codec) --------
 * Horizontal differencing of TIFF 6.0 section 14 for 8-bit samples, in place, on every strip of
 * a batch: stream i occupies data[off[i] .. off[i] + len[i]) (len == NULL: up to off[i+1]) and is
 * a sequence of rows of `row_bytes` bytes (ImageWidth * SamplesPerPixel; a short last row is
 * allowed), samples_per_pixel in 1..4.  DIFFERENCE is what a writer applies before
 * TiffStyleEncoder (lzw/src/encoder.rs:479-487) sees the strip, ACCUMULATE what a reader applies
 * to TiffStyleDecoder's output (lzw/src/decoder.rs:420-428).  The reference has no counterpart:
 * it stops at the code stream.  Device pointers, asynchronous on `cuda_stream`. */
#define SLZW_PREDICTOR_DIFFERENCE 0
#define SLZW_PREDICTOR_ACCUMULATE 1
SLZW_API int slzw_tiff_predictor_device(slzw_ctx* ctx, int direction, uint8_t* data,
                               const uint64_t* off, const uint64_t* len, uint64_t n,
                               uint32_t row_bytes, uint32_t samples_per_pixel, void* cuda_stream);
/* Makes the *_host batch entry points of this context apply the predictor on the device, inside
 * their pipeline: encode calls difference the device copy of every input stream before encoding
 * it (the caller's buffer is not modified), decode calls accumulate every decoded stream before it
 * is copied back.  (0, 0) switches it off again (the default). */
SLZW_API int slzw_set_tiff_predictor(slzw_ctx* ctx, uint32_t row_bytes, uint32_t samples_per_pixel);

/* ---- helpers ----------------------------------------------------------------------------- */
/* pinned host memory for the *_host entry points */
SLZW_API void* slzw_host_alloc(size_t bytes);
SLZW_API void slzw_host_free(void* p);
/* Formats the reference's Display text for a result (encoder.rs:31-44, decoder.rs:27-42),
 * e.g. "Code size must be between 2 and 8, was 10." ; returns bytes written (excl. NUL). */
SLZW_API int slzw_status_message(int is_decoder, uint32_t status, uint32_t detail, uint8_t code_size,
                        char* buf, size_t buf_len);

/* ---- one host batch over several GPUs of one box ----------------------------------------------
 * Streams are independent, so a batch shards by stream: contiguous ranges balanced by bytes, one
 * context and one worker thread per device, no exchange between devices; sizes and statuses are
 * gathered on the host.  Results are identical to the single-device calls (each stream is encoded /
 * decoded on its own, lzw/src/encoder.rs:273-346, decoder.rs:174-290).  The reference's API is one
 * call per job (lzw/src/lib.rs:51-91); these are that call for a box. */
typedef struct slzw_multi slzw_multi;
/* devices == NULL: the first n_devices visible devices (all of them if n_devices <= 0). */
SLZW_API int slzw_multi_create(const int* devices, int n_devices, slzw_multi** out);
SLZW_API void slzw_multi_destroy(slzw_multi* m);
SLZW_API int slzw_multi_device_count(const slzw_multi* m);
SLZW_API const char* slzw_multi_last_error(const slzw_multi* m);
SLZW_API uint64_t slzw_multi_kernel_launches(const slzw_multi* m);
/* bounds[0 .. parts]: streams [bounds[p], bounds[p+1]) go to part p; the parts are contiguous and
 * balanced by weight_off[i+1] - weight_off[i] (bytes).  The partition the calls below use. */
SLZW_API void slzw_partition_streams(const uint64_t* weight_off, uint64_t n, int parts, uint64_t* bounds);
/* slzw_encode_batch_host / slzw_decode_batch_host / slzw_encode_batch_host_dense with the batch
 * spread over the devices of `m` (encode: balanced by input bytes; decode: by the capacity slots).
 * The dense variant places the shards back to back: every device encodes and compacts its shard,
 * the host adds up the sizes, then every device copies its bytes to their final place. */
SLZW_API int slzw_multi_encode_batch_host(slzw_multi* m, const slzw_params* params, const slzw_batch* batch);
SLZW_API int slzw_multi_decode_batch_host(slzw_multi* m, const slzw_params* params, const slzw_batch* batch);
SLZW_API int slzw_multi_encode_batch_host_dense(slzw_multi* m, const slzw_params* params, const uint8_t* in,
                                       const uint64_t* in_off, uint64_t n, const uint8_t* code_size,
                                       uint64_t align, uint8_t* out_dense, uint64_t out_cap,
                                       uint64_t* out_off, uint32_t* status, uint32_t* detail,
                                       uint64_t* needed);

#ifdef __cplusplus
}
#endif
#endif /* SLZW_H */
