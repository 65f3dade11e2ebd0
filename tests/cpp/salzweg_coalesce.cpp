// Streaming facade: many threads call the reference-shaped one-stream functions of
// namespace salzweg::coalesced; the calls are merged into batched launches and every caller gets
// exactly what the plain (one launch per call) function returns -- bytes and errors.
// Usage: salzweg_coalesce <lorem_ipsum.txt> <lorem_ipsum_encoded.bin>; exits 0 when every check passes.
#include <atomic>
#include <cstdio>
#include <fstream>
#include <random>
#include <sstream>
#include <thread>

#include "salzweg.hpp"

using namespace salzweg;
using Bytes = std::vector<uint8_t>;

static std::atomic<int> failures{0};
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

// strip-like data: runs and noise, values below `limit`
static Bytes strip(uint32_t seed, size_t n, int limit) {
    std::mt19937 rng(seed);
    Bytes b(n);
    size_t i = 0;
    while (i < n) {
        const uint8_t v = (uint8_t)(rng() % limit);
        size_t run = 1 + rng() % 9;
        if (rng() % 4 == 0) run = 1;
        for (; run && i < n; run--) b[i++] = v;
    }
    return b;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const Bytes lorem = slurp(argv[1]), golden = slurp(argv[2]);
    constexpr int kThreads = 48, kPerThread = 24;

    coalesced::Options opt;
    opt.linger = std::chrono::microseconds(500);
    coalesced::set_options(opt);

    // expected results from the plain types, one launch per call
    std::vector<Bytes> raw(kThreads * kPerThread), want_tiff(raw.size()), want_gif(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
        raw[i] = strip((uint32_t)i + 1, 200 + (i * 97) % 9000, 1 << (2 + i % 7));
        if (i < 40) {
            want_tiff[i] = TiffStyleEncoder::encode_to_vec(raw[i]);
            want_gif[i] = GifStyleEncoder::encode_to_vec(raw[i], (uint8_t)(2 + i % 7));
        }
    }

    std::vector<Bytes> got_tiff(raw.size()), got_gif(raw.size()), back_tiff(raw.size()), back_gif(raw.size());
    std::vector<std::string> err_tiff(raw.size()), err_gif(raw.size());
    std::atomic<int> code_size_errors{0}, unexpected_code_errors{0}, decode_errors{0};
    auto worker = [&](int t) {
        for (int k = 0; k < kPerThread; k++) {
            const size_t i = (size_t)t * kPerThread + k;
            const uint8_t cs = (uint8_t)(2 + i % 7);
            got_tiff[i] = coalesced::TiffStyleEncoder::encode_to_vec(raw[i]);
            got_gif[i] = coalesced::GifStyleEncoder::encode_to_vec(raw[i], cs);          // per-stream code sizes
            // (the reference's decoder rejects a few of its own encoder's streams -- SURVEY F1 -- so
            // errors are recorded and compared with the plain call's, like the bytes)
            try {
                coalesced::TiffStyleDecoder::decode(got_tiff[i], back_tiff[i]);
            } catch (const DecodingError& e) {
                err_tiff[i] = e.what();
            }
            std::ostringstream os;                                                     // Write flavour
            try {
                coalesced::GifStyleDecoder::decode(got_gif[i], os, cs);
            } catch (const DecodingError& e) {
                err_gif[i] = e.what();
            }
            const std::string s = os.str();
            back_gif[i].assign(s.begin(), s.end());
            if (k == 3) {  // errors come back to the caller that caused them, neighbours are unaffected
                try {
                    coalesced::GifStyleEncoder::encode_to_vec(raw[i], 10);
                } catch (const EncodingError& e) {
                    if (e.kind == EncodingError::Kind::CodeSize && e.code_size == 10 &&
                        std::string(e.what()) == "Code size must be between 2 and 8, was 10.")
                        code_size_errors++;
                }
                try {
                    Bytes partial;
                    const Bytes bad = {1, 2, 3, 200, 1};
                    try {
                        coalesced::GifStyleEncoder::encode(bad, partial, 4);
                    } catch (const EncodingError& e) {
                        // the plain call leaves the same bytes in the writer before raising the same error
                        Bytes plain;
                        try {
                            GifStyleEncoder::encode(bad, plain, 4);
                        } catch (const EncodingError& e2) {
                            if (e.kind == EncodingError::Kind::UnexpectedCode && e.code == 200 && partial == plain &&
                                std::string(e.what()) == e2.what())
                                unexpected_code_errors++;
                        }
                    }
                } catch (...) {
                }
                try {
                    Bytes cut(got_gif[i].begin(), got_gif[i].begin() + got_gif[i].size() / 2);
                    coalesced::GifStyleDecoder::decode_to_vec(cut, cs);
                } catch (const DecodingError& e) {
                    if (e.kind == DecodingError::Kind::Io) decode_errors++;
                }
            }
        }
    };
    std::vector<std::thread> threads;
    for (int t = 0; t < kThreads; t++) threads.emplace_back(worker, t);
    for (auto& th : threads) th.join();

    size_t round_trips = 0;
    for (size_t i = 0; i < raw.size(); i++) {
        // the plain decoder's outcome for the same stream: bytes written and error text
        Bytes plain_tiff, plain_gif;
        std::string plain_err_tiff, plain_err_gif;
        try {
            TiffStyleDecoder::decode(got_tiff[i], plain_tiff);
        } catch (const DecodingError& e) {
            plain_err_tiff = e.what();
        }
        try {
            GifStyleDecoder::decode(got_gif[i], plain_gif, (uint8_t)(2 + i % 7));
        } catch (const DecodingError& e) {
            plain_err_gif = e.what();
        }
        CHECK(back_tiff[i] == plain_tiff && err_tiff[i] == plain_err_tiff);
        CHECK(back_gif[i] == plain_gif && err_gif[i] == plain_err_gif);
        if (err_tiff[i].empty()) CHECK(back_tiff[i] == raw[i]);
        if (err_gif[i].empty()) CHECK(back_gif[i] == raw[i]);
        round_trips += err_tiff[i].empty() + err_gif[i].empty();
        if (i < 40) {
            CHECK(got_tiff[i] == want_tiff[i]);
            CHECK(got_gif[i] == want_gif[i]);
        }
    }
    CHECK(round_trips * 10 > raw.size() * 2 * 9);  // the F1 rejections are rare
    CHECK(code_size_errors == kThreads);
    CHECK(unexpected_code_errors == kThreads);
    CHECK(decode_errors == kThreads);
    // the reference's golden file through the coalesced types, single caller (a batch of one)
    CHECK(coalesced::GifStyleEncoder::encode_to_vec(lorem, 7) == golden);
    CHECK(coalesced::GifStyleDecoder::decode_to_vec(golden, 7) == lorem);
    CHECK((coalesced::FixedDecoder::decode_to_vec(coalesced::FixedEncoder::encode_to_vec(lorem, Endianness::BigEndian),
                                                  Endianness::BigEndian) == lorem));

    const coalesced::Stats st = coalesced::stats();
    std::printf("coalesced: %llu streams in %llu batches (largest %llu)\n", (unsigned long long)st.streams,
                (unsigned long long)st.batches, (unsigned long long)st.largest_batch);
    CHECK(st.streams >= (uint64_t)kThreads * kPerThread * 4);
    CHECK(st.batches * 4 < st.streams);   // calls really were merged
    CHECK(st.largest_batch > 8);
    if (failures == 0) std::printf("all checks passed\n");
    return failures == 0 ? 0 : 1;
}
