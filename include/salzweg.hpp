// salzweg.hpp -- C++17 mirror of the public interface of redwarp/lzw (crate `salzweg`) over the C
// ABI of slzw.h.  Header-only; link with libslzw.so.
//
// The reference is a Rust crate and this repository's image has no Rust toolchain, so the host
// side above the C ABI is C++: the same eight types with the same associated functions, argument
// meaning and error behaviour as
//   lzw/src/lib.rs:59-91                       Endianness, CodeSizeStrategy
//   lzw/src/encoder.rs:16-52                   EncodingError (+ Display text)
//   lzw/src/decoder.rs:16-50                   DecodingError (+ Display text)
//   lzw/src/encoder.rs:199-220, 262-271        VariableEncoder::{encode, encode_to_vec}
//   lzw/src/encoder.rs:392-399, 435-439        GifStyleEncoder
//   lzw/src/encoder.rs:479-487, 519-523        TiffStyleEncoder
//   lzw/src/encoder.rs:565-576, 609-616        FixedEncoder
//   lzw/src/decoder.rs:99-120, 163-172         VariableDecoder::{decode, decode_to_vec}
//   lzw/src/decoder.rs:333-340, 378-382        GifStyleDecoder
//   lzw/src/decoder.rs:420-428, 460-464        TiffStyleDecoder
//   lzw/src/decoder.rs:503-514, 544-551        FixedDecoder
// `R: Read` becomes anything with data()/size() of bytes or a std::istream, `W: Write` a
// std::ostream or a std::vector<uint8_t>; `Result<_, E>` becomes an exception of type E thrown
// AFTER the bytes produced before the error have reached the writer, as in the reference.
// New relative to the reference: encode_batch / decode_batch (many independent streams per call) and
// namespace coalesced -- the same types for multi-threaded callers, whose concurrent one-stream calls
// are merged into batched launches.
// There is no CPU path: the first call on a machine without an sm_100 GPU throws std::runtime_error.
#ifndef SALZWEG_HPP
#define SALZWEG_HPP

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <exception>
#include <istream>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "slzw.h"

namespace salzweg {

/// lzw/src/lib.rs:59-65
enum class Endianness { BigEndian, LittleEndian };

/// lzw/src/lib.rs:71-91
enum class CodeSizeStrategy { Default, Tiff };
inline uint16_t increment(CodeSizeStrategy s) { return s == CodeSizeStrategy::Tiff ? 1 : 0; }

namespace detail {

inline std::string message(bool decoder, uint32_t status, uint32_t det, uint8_t code_size) {
    char buf[160];
    slzw_status_message(decoder ? 1 : 0, status, det, code_size, buf, sizeof buf);
    return buf;
}

// one context per thread: the reference's functions are stateless and re-entrant
inline slzw_ctx* context() {
    struct Holder {
        slzw_ctx* ctx = nullptr;
        Holder() {
            const int rc = slzw_create(0, &ctx);
            if (rc != SLZW_RC_OK)
                throw std::runtime_error(rc == SLZW_RC_NO_DEVICE
                                             ? "salzweg (B200): no usable sm_100 CUDA device, and there is no CPU path"
                                             : "salzweg (B200): slzw_create failed");
        }
        ~Holder() { slzw_destroy(ctx); }
    };
    thread_local Holder h;
    return h.ctx;
}

inline slzw_params params(uint8_t flavour, uint8_t code_size, Endianness e, CodeSizeStrategy s) {
    slzw_params p;
    p.flavour = flavour;
    p.code_size = code_size;
    p.big_endian = e == Endianness::BigEndian ? 1 : 0;
    p.tiff_early_change = s == CodeSizeStrategy::Tiff ? 1 : 0;
    return p;
}

inline std::vector<uint8_t> read_all(std::istream& in) {
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
}

struct VecWriter {
    std::vector<uint8_t>& v;
    void write_all(const uint8_t* p, size_t n) { v.insert(v.end(), p, p + n); }
};
struct StreamWriter {
    std::ostream& s;
    void write_all(const uint8_t* p, size_t n) { s.write(reinterpret_cast<const char*>(p), (std::streamsize)n); }
};

}  // namespace detail

/// lzw/src/encoder.rs:16-29; what() is the reference's Display text (encoder.rs:31-44)
struct EncodingError : std::runtime_error {
    enum class Kind { Io, CodeSize, UnexpectedCode };
    Kind kind;
    uint8_t code = 0, code_size = 0;  // CodeSize(code_size) / UnexpectedCode{code, code_size}
    uint32_t status = 0;              // slzw_status (Io: UnexpectedEof or WriteZero)
    EncodingError(Kind k, uint32_t st, uint32_t det, uint8_t cs)
        : std::runtime_error(detail::message(false, st, det, cs)), kind(k), status(st) {
        if (k == Kind::CodeSize) code_size = (uint8_t)det;
        if (k == Kind::UnexpectedCode) {
            code = (uint8_t)det;
            code_size = cs;
        }
    }
};

/// lzw/src/decoder.rs:15-25; what() is the reference's Display text (decoder.rs:27-42)
struct DecodingError : std::runtime_error {
    enum class Kind { Io, CodeSize, UnexpectedCode, MissingClearCode };
    Kind kind;
    uint16_t code = 0;     // UnexpectedCode(code)
    uint8_t code_size = 0; // CodeSize(code_size)
    uint32_t status = 0;
    DecodingError(Kind k, uint32_t st, uint32_t det)
        : std::runtime_error(detail::message(true, st, det, 0)), kind(k), status(st) {
        if (k == Kind::UnexpectedCode) code = (uint16_t)det;
        if (k == Kind::CodeSize) code_size = (uint8_t)det;
    }
};

namespace detail {

[[noreturn]] inline void launch_failure(const char* what, slzw_ctx* ctx = nullptr) {
    throw std::runtime_error(std::string(what) + ": " + slzw_last_error(ctx ? ctx : context()));
}

inline void throw_encoding(uint32_t st, uint32_t det, uint8_t cs) {
    switch (st) {
        case SLZW_ERR_CODE_SIZE: throw EncodingError(EncodingError::Kind::CodeSize, st, det, cs);
        case SLZW_ERR_UNEXPECTED_CODE: throw EncodingError(EncodingError::Kind::UnexpectedCode, st, det, cs);
        case SLZW_ERR_REFERENCE_PANIC:
            throw std::out_of_range("index out of bounds (salzweg panics on this input, encoder.rs:99)");
        default: throw EncodingError(EncodingError::Kind::Io, st, det, cs);
    }
}

inline void throw_decoding(uint32_t st, uint32_t det) {
    switch (st) {
        case SLZW_ERR_CODE_SIZE: throw DecodingError(DecodingError::Kind::CodeSize, st, det);
        case SLZW_ERR_UNEXPECTED_CODE: throw DecodingError(DecodingError::Kind::UnexpectedCode, st, det);
        case SLZW_ERR_MISSING_CLEAR_CODE: throw DecodingError(DecodingError::Kind::MissingClearCode, st, det);
        case SLZW_ERR_REFERENCE_PANIC:
            throw std::out_of_range("index out of bounds (salzweg panics on this input, decoder.rs:247)");
        default: throw DecodingError(DecodingError::Kind::Io, st, det);
    }
}

template <class W>
void run_encode(const uint8_t* data, size_t n, W into, const slzw_params& p) {
    std::vector<uint8_t> out(slzw_encode_bound(&p, n));
    uint64_t len = 0;
    uint32_t det = 0;
    const int st = slzw_encode(context(), &p, data, n, out.data(), out.size(), &len, &det);
    if (st < 0) launch_failure("slzw_encode");
    into.write_all(out.data(), (size_t)len);  // bytes produced before an error stay written
    if (st != SLZW_OK) throw_encoding((uint32_t)st, det, p.code_size);
}

template <class W>
void run_decode(const uint8_t* data, size_t n, W into, const slzw_params& p) {
    uint64_t len = 0;
    uint32_t det = 0;
    int st = slzw_decode(context(), &p, data, n, nullptr, 0, &len, &det);  // size-only pass
    if (st < 0) launch_failure("slzw_decode");
    std::vector<uint8_t> out((size_t)len ? (size_t)len : 1);
    st = slzw_decode(context(), &p, data, n, out.data(), len, &len, &det);
    if (st < 0) launch_failure("slzw_decode");
    into.write_all(out.data(), (size_t)len);
    if (st != SLZW_OK) throw_decoding((uint32_t)st, det);
}

template <bool ENC, class W>
void run_coalesced(const uint8_t* data, size_t n, W into, const slzw_params& p);

// The four codecs differ only in how their arguments map to slzw_params; ENC selects the direction,
// COALESCE whether the call is its own launch or joins the calls of other threads.
template <bool ENC, bool COALESCE, class W>
void run(const uint8_t* data, size_t n, W into, const slzw_params& p) {
    if constexpr (COALESCE) run_coalesced<ENC>(data, n, into, p);
    else if constexpr (ENC) run_encode(data, n, into, p);
    else run_decode(data, n, into, p);
}

// Argument adapters shared by every type: (bytes | istream) x (ostream | vector) and *_to_vec.
template <bool ENC, bool COALESCE, class Data, class Into>
void dispatch(Data& data, Into& into, const slzw_params& p) {
    using D = std::remove_cv_t<std::remove_reference_t<Data>>;
    using I = std::remove_cv_t<std::remove_reference_t<Into>>;
    if constexpr (std::is_base_of_v<std::istream, D>) {
        const std::vector<uint8_t> bytes = read_all(data);
        dispatch<ENC, COALESCE>(bytes, into, p);
    } else if constexpr (std::is_base_of_v<std::ostream, I>) {
        run<ENC, COALESCE>(reinterpret_cast<const uint8_t*>(data.data()), data.size(), StreamWriter{into}, p);
    } else {
        static_assert(std::is_same_v<I, std::vector<uint8_t>>, "into: std::ostream or std::vector<uint8_t>");
        run<ENC, COALESCE>(reinterpret_cast<const uint8_t*>(data.data()), data.size(), VecWriter{into}, p);
    }
}
template <bool ENC, bool COALESCE, class Data>
std::vector<uint8_t> to_vec(Data& data, const slzw_params& p) {
    std::vector<uint8_t> v;
    dispatch<ENC, COALESCE>(data, v, p);  // like the reference, the partially filled Vec is dropped with the error
    return v;
}

}  // namespace detail

// ---- encoders --------------------------------------------------------------------------------------
template <bool COALESCE>
struct BasicVariableEncoder {
    template <class R, class W>
    static void encode(R&& data, W&& into, uint8_t code_size, Endianness endianness, CodeSizeStrategy strategy) {
        detail::dispatch<true, COALESCE>(data, into, detail::params(SLZW_FLAVOUR_VARIABLE, code_size, endianness, strategy));
    }
    template <class R>
    static std::vector<uint8_t> encode_to_vec(R&& data, uint8_t code_size, Endianness endianness,
                                              CodeSizeStrategy strategy) {
        return detail::to_vec<true, COALESCE>(data, detail::params(SLZW_FLAVOUR_VARIABLE, code_size, endianness, strategy));
    }
};

template <bool COALESCE>
struct BasicGifStyleEncoder {
    template <class R, class W>
    static void encode(R&& data, W&& into, uint8_t code_size) {
        BasicVariableEncoder<COALESCE>::encode(data, into, code_size, Endianness::LittleEndian, CodeSizeStrategy::Default);
    }
    template <class R>
    static std::vector<uint8_t> encode_to_vec(R&& data, uint8_t code_size) {
        return BasicVariableEncoder<COALESCE>::encode_to_vec(data, code_size, Endianness::LittleEndian, CodeSizeStrategy::Default);
    }
};

template <bool COALESCE>
struct BasicTiffStyleEncoder {
    template <class R, class W>
    static void encode(R&& data, W&& into) {
        BasicVariableEncoder<COALESCE>::encode(data, into, 8, Endianness::BigEndian, CodeSizeStrategy::Tiff);
    }
    template <class R>
    static std::vector<uint8_t> encode_to_vec(R&& data) {
        return BasicVariableEncoder<COALESCE>::encode_to_vec(data, 8, Endianness::BigEndian, CodeSizeStrategy::Tiff);
    }
};

template <bool COALESCE>
struct BasicFixedEncoder {
    template <class R, class W>
    static void encode(R&& data, W&& into, Endianness endianness) {
        detail::dispatch<true, COALESCE>(data, into, detail::params(SLZW_FLAVOUR_FIXED, 0, endianness, CodeSizeStrategy::Default));
    }
    template <class R>
    static std::vector<uint8_t> encode_to_vec(R&& data, Endianness endianness) {
        return detail::to_vec<true, COALESCE>(data, detail::params(SLZW_FLAVOUR_FIXED, 0, endianness, CodeSizeStrategy::Default));
    }
};

// ---- decoders --------------------------------------------------------------------------------------
template <bool COALESCE>
struct BasicVariableDecoder {
    template <class R, class W>
    static void decode(R&& data, W&& into, uint8_t code_size, Endianness endianness, CodeSizeStrategy strategy) {
        detail::dispatch<false, COALESCE>(data, into, detail::params(SLZW_FLAVOUR_VARIABLE, code_size, endianness, strategy));
    }
    template <class R>
    static std::vector<uint8_t> decode_to_vec(R&& data, uint8_t code_size, Endianness endianness,
                                              CodeSizeStrategy strategy) {
        return detail::to_vec<false, COALESCE>(data, detail::params(SLZW_FLAVOUR_VARIABLE, code_size, endianness, strategy));
    }
};

template <bool COALESCE>
struct BasicGifStyleDecoder {
    template <class R, class W>
    static void decode(R&& data, W&& into, uint8_t code_size) {
        BasicVariableDecoder<COALESCE>::decode(data, into, code_size, Endianness::LittleEndian, CodeSizeStrategy::Default);
    }
    template <class R>
    static std::vector<uint8_t> decode_to_vec(R&& data, uint8_t code_size) {
        return BasicVariableDecoder<COALESCE>::decode_to_vec(data, code_size, Endianness::LittleEndian, CodeSizeStrategy::Default);
    }
};

template <bool COALESCE>
struct BasicTiffStyleDecoder {
    template <class R, class W>
    static void decode(R&& data, W&& into) {
        BasicVariableDecoder<COALESCE>::decode(data, into, 8, Endianness::BigEndian, CodeSizeStrategy::Tiff);
    }
    template <class R>
    static std::vector<uint8_t> decode_to_vec(R&& data) {
        return BasicVariableDecoder<COALESCE>::decode_to_vec(data, 8, Endianness::BigEndian, CodeSizeStrategy::Tiff);
    }
};

template <bool COALESCE>
struct BasicFixedDecoder {
    template <class R, class W>
    static void decode(R&& data, W&& into, Endianness endianness) {
        detail::dispatch<false, COALESCE>(data, into, detail::params(SLZW_FLAVOUR_FIXED, 0, endianness, CodeSizeStrategy::Default));
    }
    template <class R>
    static std::vector<uint8_t> decode_to_vec(R&& data, Endianness endianness) {
        return detail::to_vec<false, COALESCE>(data, detail::params(SLZW_FLAVOUR_FIXED, 0, endianness, CodeSizeStrategy::Default));
    }
};

/// Extension (not in the reference): VariableDecoder that tolerates deferred clear codes -- a full
/// dictionary freezes until the next clear code instead of raising MissingClearCode
/// (SLZW_FLAVOUR_VARIABLE_LENIENT in slzw.h).
template <bool COALESCE>
struct BasicLenientDecoder {
    template <class R, class W>
    static void decode(R&& data, W&& into, uint8_t code_size, Endianness endianness, CodeSizeStrategy strategy) {
        detail::dispatch<false, COALESCE>(data, into,
                                detail::params(SLZW_FLAVOUR_VARIABLE_LENIENT, code_size, endianness, strategy));
    }
    template <class R>
    static std::vector<uint8_t> decode_to_vec(R&& data, uint8_t code_size, Endianness endianness,
                                              CodeSizeStrategy strategy) {
        return detail::to_vec<false, COALESCE>(data, detail::params(SLZW_FLAVOUR_VARIABLE_LENIENT, code_size, endianness, strategy));
    }
};

// The reference's names: every call is its own launch (a batch of one).
using VariableEncoder = BasicVariableEncoder<false>;
using GifStyleEncoder = BasicGifStyleEncoder<false>;
using TiffStyleEncoder = BasicTiffStyleEncoder<false>;
using FixedEncoder = BasicFixedEncoder<false>;
using VariableDecoder = BasicVariableDecoder<false>;
using GifStyleDecoder = BasicGifStyleDecoder<false>;
using TiffStyleDecoder = BasicTiffStyleDecoder<false>;
using FixedDecoder = BasicFixedDecoder<false>;
using LenientDecoder = BasicLenientDecoder<false>;

/// Streaming facade.  The same types and signatures, for callers that keep the
/// reference's one-stream-per-call shape but call from many threads (a rayon-style parallel loop
/// over strips or frames): concurrent calls with the same flavour / bit order are coalesced into
/// one batched launch (group commit), each caller gets its own stream's bytes and error back.
namespace coalesced {
using VariableEncoder = BasicVariableEncoder<true>;
using GifStyleEncoder = BasicGifStyleEncoder<true>;
using TiffStyleEncoder = BasicTiffStyleEncoder<true>;
using FixedEncoder = BasicFixedEncoder<true>;
using VariableDecoder = BasicVariableDecoder<true>;
using GifStyleDecoder = BasicGifStyleDecoder<true>;
using TiffStyleDecoder = BasicTiffStyleDecoder<true>;
using FixedDecoder = BasicFixedDecoder<true>;
using LenientDecoder = BasicLenientDecoder<true>;
}  // namespace coalesced

// ---- batches (new relative to the reference) --------------------------------------------------------
/// One stream's outcome of a batched call: the bytes produced (also those before an error) and the
/// status / detail of include/slzw.h (0 = ok).
struct StreamResult {
    std::vector<uint8_t> bytes;
    uint32_t status = 0, detail = 0;
    bool ok() const { return status == SLZW_OK; }
};

namespace detail {

inline std::vector<StreamResult> collect(const std::vector<uint8_t>& out, const std::vector<uint64_t>& out_off,
                                         const std::vector<uint64_t>& len, const std::vector<uint32_t>& st,
                                         const std::vector<uint32_t>& det) {
    std::vector<StreamResult> res(len.size());
    for (size_t i = 0; i < res.size(); i++) {
        res[i].bytes.assign(out.begin() + out_off[i], out.begin() + out_off[i] + len[i]);
        res[i].status = st[i];
        res[i].detail = det[i];
    }
    return res;
}

inline std::vector<StreamResult> encode_batch(slzw_ctx* ctx, const slzw_params& p, const uint8_t* data,
                                              const std::vector<uint64_t>& offsets, const uint8_t* code_sizes) {
    const uint64_t n = offsets.empty() ? 0 : offsets.size() - 1;
    if (n == 0) return {};
    std::vector<uint64_t> out_off(n + 1, 0), len(n);
    for (uint64_t i = 0; i < n; i++)
        out_off[i + 1] = out_off[i] + ((slzw_encode_bound(&p, offsets[i + 1] - offsets[i]) + 15) & ~15ull);
    std::vector<uint8_t> out(out_off[n] ? out_off[n] : 1);
    std::vector<uint32_t> st(n), det(n);
    slzw_batch b{data, offsets.data(), out.data(), out_off.data(), len.data(), st.data(), det.data(), code_sizes, n};
    if (slzw_encode_batch_host(ctx, &p, &b) != SLZW_RC_OK) launch_failure("slzw_encode_batch_host", ctx);
    return collect(out, out_off, len, st, det);
}

// capacities == nullptr: a size-only pass first, as decode_to_vec has no caller-supplied capacity
inline std::vector<StreamResult> decode_batch(slzw_ctx* ctx, const slzw_params& p, const uint8_t* data,
                                              const std::vector<uint64_t>& offsets, const uint64_t* capacities,
                                              const uint8_t* code_sizes) {
    const uint64_t n = offsets.empty() ? 0 : offsets.size() - 1;
    if (n == 0) return {};
    std::vector<uint64_t> out_off(n + 1, 0), len(n);
    std::vector<uint32_t> st(n), det(n);
    slzw_batch b{data, offsets.data(), nullptr, nullptr, len.data(), st.data(), det.data(), code_sizes, n};
    if (!capacities) {
        if (slzw_decoded_sizes_batch_host(ctx, &p, &b) != SLZW_RC_OK)
            launch_failure("slzw_decoded_sizes_batch_host", ctx);
        capacities = len.data();
    }
    for (uint64_t i = 0; i < n; i++) out_off[i + 1] = out_off[i] + capacities[i];
    std::vector<uint8_t> out(out_off[n] ? out_off[n] : 1);
    b.out = out.data();
    b.out_off = out_off.data();
    if (slzw_decode_batch_host(ctx, &p, &b) != SLZW_RC_OK) launch_failure("slzw_decode_batch_host", ctx);
    return collect(out, out_off, len, st, det);
}

}  // namespace detail

/// Encodes streams data[offsets[i] .. offsets[i+1]) with one configuration; per-stream code sizes
/// (GIF frames with different palettes) may be given in `code_sizes`.
inline std::vector<StreamResult> encode_batch(const uint8_t* data, const std::vector<uint64_t>& offsets,
                                              uint8_t flavour, uint8_t code_size, Endianness endianness,
                                              CodeSizeStrategy strategy,
                                              const std::vector<uint8_t>* code_sizes = nullptr) {
    return detail::encode_batch(detail::context(), detail::params(flavour, code_size, endianness, strategy), data, offsets,
                                code_sizes ? code_sizes->data() : nullptr);
}

/// Decodes streams data[offsets[i] .. offsets[i+1]); capacities[i] is the room given to stream i
/// (the analogue of a `&mut [u8]` writer, e.g. the strip size of a TIFF).
inline std::vector<StreamResult> decode_batch(const uint8_t* data, const std::vector<uint64_t>& offsets,
                                              const std::vector<uint64_t>& capacities, uint8_t flavour,
                                              uint8_t code_size, Endianness endianness, CodeSizeStrategy strategy,
                                              const std::vector<uint8_t>* code_sizes = nullptr) {
    return detail::decode_batch(detail::context(), detail::params(flavour, code_size, endianness, strategy), data, offsets,
                                capacities.data(), code_sizes ? code_sizes->data() : nullptr);
}

// ---- the same batches over every GPU of the box (slzw_multi_*) -------------------------------------
/// `salzweg::multi::encode_batch` / `decode_batch`: identical signatures and results; the batch is
/// sharded by bytes over all visible sm_100 devices (one context and one worker thread each, no
/// exchange between devices, sizes gathered on the host).  The device set is opened on first use.
namespace multi {
namespace detail {
inline slzw_multi* box() {
    static slzw_multi* m = [] {
        slzw_multi* x = nullptr;
        const int rc = slzw_multi_create(nullptr, 0, &x);
        if (rc != SLZW_RC_OK)
            throw std::runtime_error(rc == SLZW_RC_NO_DEVICE ? "salzweg: no sm_100 CUDA device (there is no CPU fallback)"
                                                              : "salzweg: slzw_multi_create failed");
        return x;
    }();
    return m;
}
[[noreturn]] inline void failure(const char* what) {
    throw std::runtime_error(std::string("salzweg: ") + what + " failed: " + slzw_multi_last_error(box()));
}
}  // namespace detail

inline int device_count() { return slzw_multi_device_count(detail::box()); }

inline std::vector<StreamResult> encode_batch(const uint8_t* data, const std::vector<uint64_t>& offsets,
                                              uint8_t flavour, uint8_t code_size, Endianness endianness,
                                              CodeSizeStrategy strategy,
                                              const std::vector<uint8_t>* code_sizes = nullptr) {
    const uint64_t n = offsets.empty() ? 0 : offsets.size() - 1;
    if (n == 0) return {};
    const slzw_params p = salzweg::detail::params(flavour, code_size, endianness, strategy);
    std::vector<uint64_t> out_off(n + 1, 0), len(n);
    for (uint64_t i = 0; i < n; i++)
        out_off[i + 1] = out_off[i] + ((slzw_encode_bound(&p, offsets[i + 1] - offsets[i]) + 15) & ~15ull);
    std::vector<uint8_t> out(out_off[n] ? out_off[n] : 1);
    std::vector<uint32_t> st(n), det(n);
    slzw_batch b{data, offsets.data(), out.data(), out_off.data(), len.data(), st.data(), det.data(),
                 code_sizes ? code_sizes->data() : nullptr, n};
    if (slzw_multi_encode_batch_host(detail::box(), &p, &b) != SLZW_RC_OK) detail::failure("slzw_multi_encode_batch_host");
    return salzweg::detail::collect(out, out_off, len, st, det);
}

inline std::vector<StreamResult> decode_batch(const uint8_t* data, const std::vector<uint64_t>& offsets,
                                              const std::vector<uint64_t>& capacities, uint8_t flavour,
                                              uint8_t code_size, Endianness endianness, CodeSizeStrategy strategy,
                                              const std::vector<uint8_t>* code_sizes = nullptr) {
    const uint64_t n = offsets.empty() ? 0 : offsets.size() - 1;
    if (n == 0) return {};
    const slzw_params p = salzweg::detail::params(flavour, code_size, endianness, strategy);
    std::vector<uint64_t> out_off(n + 1, 0), len(n);
    for (uint64_t i = 0; i < n; i++) out_off[i + 1] = out_off[i] + capacities[i];
    std::vector<uint8_t> out(out_off[n] ? out_off[n] : 1);
    std::vector<uint32_t> st(n), det(n);
    slzw_batch b{data, offsets.data(), out.data(), out_off.data(), len.data(), st.data(), det.data(),
                 code_sizes ? code_sizes->data() : nullptr, n};
    if (slzw_multi_decode_batch_host(detail::box(), &p, &b) != SLZW_RC_OK) detail::failure("slzw_multi_decode_batch_host");
    return salzweg::detail::collect(out, out_off, len, st, det);
}
}  // namespace multi

// ---- coalescing of one-stream calls -------------------------------------------------
namespace coalesced {

/// How long the thread that runs a batch waits for other callers to join it, and when it stops
/// waiting early.  Process-wide; the defaults suit a parallel loop over strips or frames.
struct Options {
    std::chrono::microseconds linger{200};
    size_t max_streams = 65536;
    size_t max_bytes = size_t(256) << 20;
};
struct Stats {
    uint64_t streams = 0, batches = 0, largest_batch = 0;
};

}  // namespace coalesced

namespace detail {

// Group commit: a caller queues its stream; if nobody is running a batch it becomes the leader,
// lingers a moment, takes everything queued (its own stream included), runs ONE batched call on its
// the coalescer's own context and hands the results out.  Streams that arrive while a batch runs wait for
// the next leader.  One instance per (direction, flavour, bit order, width strategy): the code
// size travels per stream.
class Coalescer {
  public:
    Coalescer(bool encoder, slzw_params p) : encoder_(encoder), params_(p) {}
    ~Coalescer() { slzw_destroy(ctx_); }
    Coalescer(const Coalescer&) = delete;
    Coalescer& operator=(const Coalescer&) = delete;

    StreamResult run(const uint8_t* data, size_t n, uint8_t code_size) {
        Request r;
        r.data = data, r.n = n, r.code_size = code_size;
        std::unique_lock<std::mutex> lk(mu_);
        queue_.push_back(&r);
        queued_bytes_ += n;
        if (leader_ && (queue_.size() >= options_.max_streams || queued_bytes_ >= options_.max_bytes)) cv_.notify_all();
        while (!r.done) {
            if (leader_) {
                cv_.wait(lk);
                continue;
            }
            leader_ = true;
            const auto deadline = std::chrono::steady_clock::now() + options_.linger;
            while (queue_.size() < options_.max_streams && queued_bytes_ < options_.max_bytes &&
                   cv_.wait_until(lk, deadline) != std::cv_status::timeout) {
            }
            std::vector<Request*> batch;
            batch.swap(queue_);
            queued_bytes_ = 0;
            lk.unlock();
            std::exception_ptr failure;
            try {
                execute(batch);
            } catch (...) {
                failure = std::current_exception();  // a failed launch fails every stream of the batch
            }
            lk.lock();
            for (Request* q : batch) {
                q->failure = failure;
                q->done = true;
            }
            stats_.streams += batch.size();
            stats_.batches += 1;
            if (batch.size() > stats_.largest_batch) stats_.largest_batch = batch.size();
            leader_ = false;
            cv_.notify_all();
        }
        lk.unlock();
        if (r.failure) std::rethrow_exception(r.failure);
        return std::move(r.result);
    }

    void set_options(const coalesced::Options& o) {
        std::lock_guard<std::mutex> lk(mu_);
        options_ = o;
    }
    coalesced::Stats stats() {
        std::lock_guard<std::mutex> lk(mu_);
        return stats_;
    }

  private:
    struct Request {
        const uint8_t* data = nullptr;
        size_t n = 0;
        uint8_t code_size = 0;
        bool done = false;
        StreamResult result;
        std::exception_ptr failure;
    };

    // only ever entered by the current leader
    void execute(const std::vector<Request*>& batch) {
        if (!ctx_) {
            const int rc = slzw_create(0, &ctx_);
            if (rc != SLZW_RC_OK)
                throw std::runtime_error(rc == SLZW_RC_NO_DEVICE
                                             ? "salzweg (B200): no usable sm_100 CUDA device, and there is no CPU path"
                                             : "salzweg (B200): slzw_create failed");
        }
        std::vector<uint64_t> offsets(batch.size() + 1, 0);
        for (size_t i = 0; i < batch.size(); i++) offsets[i + 1] = offsets[i] + batch[i]->n;
        std::vector<uint8_t> data(offsets.back() ? offsets.back() : 1), code_sizes(batch.size());
        for (size_t i = 0; i < batch.size(); i++) {
            if (batch[i]->n) std::memcpy(data.data() + offsets[i], batch[i]->data, batch[i]->n);
            code_sizes[i] = batch[i]->code_size;
        }
        std::vector<StreamResult> res =
            encoder_ ? encode_batch(ctx_, params_, data.data(), offsets, code_sizes.data())
                     : decode_batch(ctx_, params_, data.data(), offsets, nullptr, code_sizes.data());
        for (size_t i = 0; i < batch.size(); i++) batch[i]->result = std::move(res[i]);
    }

    const bool encoder_;
    const slzw_params params_;
    slzw_ctx* ctx_ = nullptr;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Request*> queue_;
    size_t queued_bytes_ = 0;
    bool leader_ = false;
    coalesced::Options options_;
    coalesced::Stats stats_;
};

struct CoalescerRegistry {
    std::mutex mu;
    std::map<uint32_t, std::unique_ptr<Coalescer>> by_key;
    coalesced::Options options;
    static CoalescerRegistry& instance() {
        static CoalescerRegistry r;
        return r;
    }
    Coalescer& get(bool encoder, const slzw_params& p) {
        const uint32_t key = (encoder ? 1u : 0u) | (uint32_t)p.flavour << 8 | (uint32_t)p.big_endian << 16 |
                             (uint32_t)p.tiff_early_change << 24;
        std::lock_guard<std::mutex> lk(mu);
        std::unique_ptr<Coalescer>& c = by_key[key];
        if (!c) {
            c.reset(new Coalescer(encoder, p));
            c->set_options(options);
        }
        return *c;
    }
};

template <bool ENC, class W>
void run_coalesced(const uint8_t* data, size_t n, W into, const slzw_params& p) {
    StreamResult r = CoalescerRegistry::instance().get(ENC, p).run(data, n, p.code_size);
    into.write_all(r.bytes.data(), r.bytes.size());  // bytes produced before an error stay written
    if (!r.ok()) {
        if (ENC) throw_encoding(r.status, r.detail, p.code_size);
        else throw_decoding(r.status, r.detail);
    }
}

}  // namespace detail

namespace coalesced {

inline void set_options(const Options& o) {
    detail::CoalescerRegistry& r = detail::CoalescerRegistry::instance();
    std::lock_guard<std::mutex> lk(r.mu);
    r.options = o;
    for (auto& kv : r.by_key) kv.second->set_options(o);
}

/// Totals over every coalescer of the process (streams / batches = average batch size).
inline Stats stats() {
    detail::CoalescerRegistry& r = detail::CoalescerRegistry::instance();
    std::lock_guard<std::mutex> lk(r.mu);
    Stats t;
    for (auto& kv : r.by_key) {
        const Stats s = kv.second->stats();
        t.streams += s.streams;
        t.batches += s.batches;
        if (s.largest_batch > t.largest_batch) t.largest_batch = s.largest_batch;
    }
    return t;
}

}  // namespace coalesced

}  // namespace salzweg

#endif  // SALZWEG_HPP
