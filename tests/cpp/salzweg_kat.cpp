// The reference's own unit tests and doctests (lzw/src/encoder.rs:665-835, lzw/src/decoder.rs:649-769,
// lzw/src/lib.rs:13-49), restated against include/salzweg.hpp.  Usage: salzweg_kat <lorem_ipsum.txt>
// <lorem_ipsum_encoded.bin>; exits 0 when every check passes.
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

#include "salzweg.hpp"

using namespace salzweg;
using Bytes = std::vector<uint8_t>;

static int failures = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                 \
        }                                                               \
    } while (0)

static Bytes slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return Bytes(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const Bytes lorem = slurp(argv[1]), golden = slurp(argv[2]);
    const Bytes d4 = {0, 0, 1, 3};
    const Bytes d40 = {1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2,
                       1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2};

    // lib.rs:21-31, encoder.rs:689-712, 816-835
    CHECK((GifStyleEncoder::encode_to_vec(d4, 2) == Bytes{0x04, 0x32, 0x05}));
    CHECK((TiffStyleEncoder::encode_to_vec(d4) == Bytes{0x80, 0x00, 0x00, 0x00, 0x10, 0x1c, 0x04}));
    CHECK((FixedEncoder::encode_to_vec(d4, Endianness::LittleEndian) == Bytes{0x00, 0x00, 0x00, 0x01, 0x30, 0x00}));
    CHECK((VariableEncoder::encode_to_vec(d4, 2, Endianness::LittleEndian, CodeSizeStrategy::Default) ==
           Bytes{0x04, 0x32, 0x05}));
    // encoder.rs:666-686 / decoder.rs:650-672
    const Bytes e40 = {0x8C, 0x2D, 0x99, 0x87, 0x2A, 0x1C, 0xDC, 0x33, 0xA0, 0x02, 0x55, 0x00};
    CHECK(GifStyleEncoder::encode_to_vec(d40, 2) == e40);
    CHECK(GifStyleDecoder::decode_to_vec(e40, 2) == d40);
    CHECK(TiffStyleDecoder::decode_to_vec(TiffStyleEncoder::encode_to_vec(d40)) == d40);
    CHECK(FixedDecoder::decode_to_vec(FixedEncoder::encode_to_vec(d40, Endianness::BigEndian), Endianness::BigEndian) == d40);
    // encoder.rs:740-755 / decoder.rs:703-718: the golden file
    CHECK(GifStyleEncoder::encode_to_vec(lorem, 7) == golden);
    CHECK(GifStyleDecoder::decode_to_vec(golden, 7) == lorem);
    // Read / Write flavours of the same calls
    {
        std::istringstream in(std::string(lorem.begin(), lorem.end()));
        std::ostringstream out;
        GifStyleEncoder::encode(in, out, 7);
        const std::string s = out.str();
        CHECK(Bytes(s.begin(), s.end()) == golden);
        Bytes back;
        GifStyleDecoder::decode(golden, back, 7);
        CHECK(back == lorem);
    }
    // encoder.rs:758-774, decoder.rs:721-737: Display texts differ by the trailing period
    try {
        GifStyleEncoder::encode_to_vec(d4, 10);
        CHECK(false);
    } catch (const EncodingError& e) {
        CHECK(e.kind == EncodingError::Kind::CodeSize && e.code_size == 10);
        CHECK(std::string(e.what()) == "Code size must be between 2 and 8, was 10.");
    }
    try {
        GifStyleDecoder::decode_to_vec(d4, 10);
        CHECK(false);
    } catch (const DecodingError& e) {
        CHECK(e.kind == DecodingError::Kind::CodeSize && e.code_size == 10);
        CHECK(std::string(e.what()) == "Code size must be between 2 and 8, was 10");
    }
    // encoder.rs:777-795
    try {
        GifStyleEncoder::encode_to_vec(Bytes{0, 1, 8, 3}, 2);
        CHECK(false);
    } catch (const EncodingError& e) {
        CHECK(e.kind == EncodingError::Kind::UnexpectedCode && e.code == 8 && e.code_size == 2);
        CHECK(std::string(e.what()) == "Unexpected code 8. For code size 2, data should be < 4.");
    }
    // decoder.rs:759-769
    try {
        TiffStyleDecoder::decode_to_vec(Bytes{0x1F, 0x40, 0x3A, 0x00, 0x00, 0x00, 0x44, 0x00, 0x00, 0x44, 0x00, 0x60, 0x54});
        CHECK(false);
    } catch (const DecodingError& e) {
        CHECK(e.kind == DecodingError::Kind::UnexpectedCode && e.code == 258);
        CHECK(std::string(e.what()) == "Unexpected code while decompressing: 258");
    }
    // bytes written before an error stay in the writer (encoder.rs:315-317 returns after emitting)
    {
        Bytes partial;
        try {
            GifStyleEncoder::encode(Bytes{0, 1, 2, 3, 0, 1, 2, 3, 9}, partial, 2);
            CHECK(false);
        } catch (const EncodingError&) {
            CHECK(!partial.empty());
        }
    }
    // batches: every stream encodes as it would on its own
    {
        Bytes all = lorem;
        all.insert(all.end(), d40.begin(), d40.end());
        const std::vector<uint64_t> off = {0, lorem.size(), lorem.size(), all.size()};  // incl. an empty stream
        auto enc = encode_batch(all.data(), off, SLZW_FLAVOUR_VARIABLE, 8, Endianness::BigEndian, CodeSizeStrategy::Tiff);
        CHECK(enc.size() == 3 && enc[0].ok() && enc[1].ok() && enc[2].ok());
        CHECK(enc[0].bytes == TiffStyleEncoder::encode_to_vec(lorem));
        CHECK((enc[1].bytes == Bytes{0x80, 0x40, 0x40}));
        CHECK(enc[2].bytes == TiffStyleEncoder::encode_to_vec(d40));
        Bytes packed;
        std::vector<uint64_t> poff = {0};
        for (auto& r : enc) {
            packed.insert(packed.end(), r.bytes.begin(), r.bytes.end());
            poff.push_back(packed.size());
        }
        auto dec = decode_batch(packed.data(), poff, {lorem.size(), 0, d40.size()}, SLZW_FLAVOUR_VARIABLE, 8,
                                Endianness::BigEndian, CodeSizeStrategy::Tiff);
        CHECK(dec[0].bytes == lorem && dec[1].bytes.empty() && dec[2].bytes == d40);
        // the same batch over every GPU of the box: identical results
        auto menc = multi::encode_batch(all.data(), off, SLZW_FLAVOUR_VARIABLE, 8, Endianness::BigEndian,
                                        CodeSizeStrategy::Tiff);
        CHECK(multi::device_count() >= 1 && menc.size() == 3);
        for (size_t i = 0; i < menc.size() && i < enc.size(); i++)
            CHECK(menc[i].bytes == enc[i].bytes && menc[i].status == enc[i].status);
        auto mdec = multi::decode_batch(packed.data(), poff, {lorem.size(), 0, d40.size()}, SLZW_FLAVOUR_VARIABLE, 8,
                                        Endianness::BigEndian, CodeSizeStrategy::Tiff);
        CHECK(mdec[0].bytes == lorem && mdec[1].bytes.empty() && mdec[2].bytes == d40);
    }
    std::printf(failures ? "%d check(s) failed\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
