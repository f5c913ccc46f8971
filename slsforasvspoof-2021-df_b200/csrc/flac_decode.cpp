// Native FLAC decoder for the audio ingest: the ASVspoof corpora are 16 kHz / 16-bit mono FLAC,
// which the reference decodes one file per Dataset.__getitem__ through librosa -> libsndfile (data_utils_SSL.py:109-113).
// Written from the format specification (RFC 9639): STREAMINFO, frame headers (CRC-8), CONSTANT / VERBATIM / FIXED / LPC
// subframes, partitioned Rice residuals incl. escape partitions, wasted bits, the four stereo decorrelation modes, frame
// CRC-16 and the STREAMINFO MD5 of the decoded PCM.  Known answers: the three worked examples of RFC 9639 appendix D
// (tests/test_host.py).  Host code only (no CUDA): the device side starts at 16-bit PCM (ingest_pcm16_kernel).
//
// pad() keeps only the first 64 600 samples of a clip (data_utils_SSL.py:60-61), so the decoder stops after `max_samples`
// samples per channel - a 10-second utterance costs the decode of its first 4 seconds.
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/slsb200.h"

namespace {

enum {
    FLAC_OK = 0, FLAC_E_MAGIC = -2, FLAC_E_TRUNC = -3, FLAC_E_HEADER = -4, FLAC_E_CRC8 = -5, FLAC_E_CRC16 = -6, FLAC_E_RESERVED = -7,
    FLAC_E_MD5 = -8, FLAC_E_UNSUPPORTED = -9, FLAC_E_STREAMINFO = -10, FLAC_E_ARG = -11
};

// MSB-first bit reader over a byte buffer with a 64-bit window (refilled a byte at a time, at most 8 loads per refill).
struct BitReader {
    const uint8_t* p; int64_t n; int64_t next = 0;   // next byte to load into the window
    uint64_t win = 0; int cnt = 0;                   // `cnt` valid bits at the top of `win`
    bool fail = false;
    BitReader(const uint8_t* d, int64_t bytes) : p(d), n(bytes) {}
    inline void refill() {
        if (next + 8 <= n) {                                   // one unaligned 8-byte load, whole bytes that fit are consumed
            uint64_t v;
            memcpy(&v, p + next, 8);
            v = __builtin_bswap64(v);
            const int take = (64 - cnt) >> 3;                  // bytes that fit above the `cnt` valid bits
            if (take == 8) { win = v; }
            else if (take > 0) { win |= (v >> (64 - 8 * take)) << (64 - cnt - 8 * take); }
            next += take; cnt += 8 * take;
            return;
        }
        while (cnt <= 56 && next < n) { win |= (uint64_t)p[next++] << (56 - cnt); cnt += 8; }
    }
    inline uint32_t bits(int k) {                   // k in 0..32
        if (k == 0) return 0;
        if (cnt < k) { refill(); if (cnt < k) { fail = true; cnt = 0; win = 0; return 0; } }
        const uint32_t v = (uint32_t)(win >> (64 - k));
        win <<= k; cnt -= k;
        return v;
    }
    inline int32_t sbits(int k) {                   // two's complement, k in 1..32
        const uint32_t v = bits(k);
        if (k == 32) return (int32_t)v;
        const uint32_t m = 1u << (k - 1);
        return (int32_t)((v ^ m) - m);
    }
    inline uint32_t unary() {                       // number of 0 bits before the next 1 bit (which is consumed too)
        uint32_t z = 0;
        for (;;) {
            if (cnt == 0) { refill(); if (cnt == 0) { fail = true; return 0; } }
            if (win == 0) { z += cnt; cnt = 0; continue; }          // only zeros left in the window (bits below `cnt` are zero by construction)
            const int lead = __builtin_clzll(win);
            z += lead;
            win = lead == 63 ? 0 : win << (lead + 1);
            cnt -= lead + 1;
            return z;
        }
    }
    inline int64_t bitpos() const { return next * 8 - cnt; }
    inline void align() { const int r = cnt & 7; win <<= r; cnt -= r; }   // `next * 8` is byte aligned, so cnt mod 8 = bits left in the current byte
};

uint8_t crc8_table[256]; uint16_t crc16_table[256]; uint16_t crc16_slice[8][256];
// Built exactly once, race-free: the decoder runs on many threads at once (ingest.py's thread pool calls it with the GIL
// released), and a function-local static is initialised under the C++11 guarantee (one thread builds, the others wait).
void build_tables() {
    for (int i = 0; i < 256; ++i) {
        uint8_t c = (uint8_t)i; uint16_t d = (uint16_t)(i << 8);
        for (int b = 0; b < 8; ++b) { c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1)); d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1)); }
        crc8_table[i] = c; crc16_table[i] = d;
    }
    // slicing-by-8: slice[k][b] = CRC of byte b followed by k zero bytes
    for (int i = 0; i < 256; ++i) {
        uint16_t c = crc16_table[i];
        crc16_slice[0][i] = c;
        for (int k = 1; k < 8; ++k) { c = (uint16_t)((c << 8) ^ crc16_table[c >> 8]); crc16_slice[k][i] = c; }
    }
}
void init_tables() {
    static const bool once = (build_tables(), true);
    (void)once;
}
uint8_t crc8(const uint8_t* p, int64_t n) { uint8_t c = 0; for (int64_t i = 0; i < n; ++i) c = crc8_table[c ^ p[i]]; return c; }
uint16_t crc16_update(uint16_t c, const uint8_t* p, int64_t n) {
    int64_t i = 0;
    for (; i + 8 <= n; i += 8)
        c = (uint16_t)(crc16_slice[7][p[i] ^ (c >> 8)] ^ crc16_slice[6][p[i + 1] ^ (c & 0xFF)] ^ crc16_slice[5][p[i + 2]] ^ crc16_slice[4][p[i + 3]] ^
                       crc16_slice[3][p[i + 4]] ^ crc16_slice[2][p[i + 5]] ^ crc16_slice[1][p[i + 6]] ^ crc16_slice[0][p[i + 7]]);
    for (; i < n; ++i) c = (uint16_t)((c << 8) ^ crc16_table[(c >> 8) ^ p[i]]);
    return c;
}
uint16_t crc16(const uint8_t* p, int64_t n) { return crc16_update(0, p, n); }

// ---- MD5 (RFC 1321) of the interleaved little-endian PCM, as STREAMINFO stores it ----
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u; uint64_t len = 0; uint8_t buf[64]; int fill = 0;
    static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
    void block(const uint8_t* m) {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be,
            0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
            0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
            0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t w[16];
        for (int i = 0; i < 16; ++i) w[i] = (uint32_t)m[4 * i] | ((uint32_t)m[4 * i + 1] << 8) | ((uint32_t)m[4 * i + 2] << 16) | ((uint32_t)m[4 * i + 3] << 24);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; ++i) {
            uint32_t f; int g;
            if (i < 16) { f = (B & C) | (~B & D); g = i; }
            else if (i < 32) { f = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = B ^ C ^ D; g = (3 * i + 5) & 15; }
            else { f = C ^ (B | ~D); g = (7 * i) & 15; }
            const uint32_t t = D; D = C; C = B; B = B + rol(A + f + K[i] + w[g], S[i]); A = t;
        }
        a += A; b += B; c += C; d += D;
    }
    void update(const uint8_t* p, size_t n) {
        len += n;
        if (fill == 0) { while (n >= 64) { block(p); p += 64; n -= 64; } }
        while (n > 0) {
            const size_t take = (size_t)(64 - fill) < n ? (size_t)(64 - fill) : n;
            memcpy(buf + fill, p, take); fill += (int)take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    void final(uint8_t out[16]) {
        const uint64_t bitlen = len * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56) update(&zero, 1);
        uint8_t l[8];
        for (int i = 0; i < 8; ++i) l[i] = (uint8_t)(bitlen >> (8 * i));
        update(l, 8);
        const uint32_t v[4] = {a, b, c, d};
        for (int i = 0; i < 16; ++i) out[i] = (uint8_t)(v[i >> 2] >> (8 * (i & 3)));
    }
};

struct StreamInfo { int min_block = 0, max_block = 0, rate = 0, channels = 0, bps = 0; int64_t total = 0; uint8_t md5[16] = {0}; };

int read_residual(BitReader& br, int32_t* res, int blocksize, int order) {
    const uint32_t method = br.bits(2);
    if (method > 1) return FLAC_E_RESERVED;
    const int pbits = method == 0 ? 4 : 5, esc = method == 0 ? 15 : 31;
    const int porder = (int)br.bits(4);
    const int parts = 1 << porder;
    if ((blocksize >> porder) << porder != blocksize && porder > 0) return FLAC_E_HEADER;
    int i = 0;
    for (int p = 0; p < parts; ++p) {
        int count = (blocksize >> porder) - (p == 0 ? order : 0);
        if (count < 0) return FLAC_E_HEADER;
        const int k = (int)br.bits(pbits);
        if (k == esc) {
            const int raw = (int)br.bits(5);
            for (int j = 0; j < count; ++j) res[i++] = raw ? br.sbits(raw) : 0;
        } else {
            // bit-reader state in locals for the hot loop (the int32 stores to `res` may alias the reader's members otherwise)
            uint64_t win = br.win; int cnt = br.cnt; int64_t next = br.next;
            const uint8_t* const p = br.p; const int64_t n = br.n;
            for (int j = 0; j < count; ++j) {
                if (cnt <= 56 && next + 8 <= n) {                          // 8-byte refill, whole bytes only (taken on nearly every sample: predictable)
                    uint64_t v;
                    memcpy(&v, p + next, 8);
                    v = __builtin_bswap64(v);
                    const int take = (64 - cnt) >> 3;
                    win |= (v >> (64 - 8 * take)) << (64 - cnt - 8 * take);   // cnt in 0..56 here, so take is 1..8 and both shifts are < 64
                    next += take; cnt += 8 * take;
                }
                uint32_t u;
                const int lead = win ? __builtin_clzll(win) : 64;
                if (lead + 1 + k <= cnt && lead + 1 + k < 64) {            // whole code word inside the window: one step
                    const uint64_t w = win << (lead + 1);
                    u = ((uint32_t)lead << k) | (k ? (uint32_t)(w >> (64 - k)) : 0u);
                    win = w << k; cnt -= lead + 1 + k;
                } else {                                                    // long code word or end of data: the careful path
                    br.win = win; br.cnt = cnt; br.next = next;
                    const uint32_t q = br.unary();
                    u = (q << k) | (k ? br.bits(k) : 0u);
                    win = br.win; cnt = br.cnt; next = br.next;
                }
                res[i++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);           // zig-zag: even -> u / 2, odd -> -(u + 1) / 2
            }
            br.win = win; br.cnt = cnt; br.next = next;
        }
        if (br.fail) return FLAC_E_TRUNC;
    }
    return FLAC_OK;
}

// out[i] = res[i] + (sum_j coef[j] * out[i - 1 - j] >> shift); unsigned accumulation: corrupt streams may wrap, valid ones never get
// near 2^63.  The common orders are compiled with the tap loop unrolled.
template <int ORDER>
inline void predict_n(int64_t* out, const int32_t* res, const int32_t* coef, int order, int shift, int blocksize) {
    const int n = ORDER > 0 ? ORDER : order;
    int64_t c[ORDER > 0 ? ORDER : 32];
    for (int j = 0; j < n; ++j) c[j] = coef[j];
    for (int i = n; i < blocksize; ++i) {
        uint64_t acc = 0;
        for (int j = 0; j < n; ++j) acc += (uint64_t)c[j] * (uint64_t)out[i - 1 - j];
        out[i] = (int64_t)((uint64_t)((int64_t)acc >> shift) + (uint64_t)(int64_t)res[i]);
    }
}
void predict(int64_t* out, const int32_t* res, const int32_t* coef, int order, int shift, int blocksize) {
    switch (order) {
        case 0: for (int i = 0; i < blocksize; ++i) out[i] = res[i]; break;
        case 1: predict_n<1>(out, res, coef, order, shift, blocksize); break;
        case 2: predict_n<2>(out, res, coef, order, shift, blocksize); break;
        case 3: predict_n<3>(out, res, coef, order, shift, blocksize); break;
        case 4: predict_n<4>(out, res, coef, order, shift, blocksize); break;
        case 5: predict_n<5>(out, res, coef, order, shift, blocksize); break;
        case 6: predict_n<6>(out, res, coef, order, shift, blocksize); break;
        case 7: predict_n<7>(out, res, coef, order, shift, blocksize); break;
        case 8: predict_n<8>(out, res, coef, order, shift, blocksize); break;
        case 9: predict_n<9>(out, res, coef, order, shift, blocksize); break;
        case 10: predict_n<10>(out, res, coef, order, shift, blocksize); break;
        case 11: predict_n<11>(out, res, coef, order, shift, blocksize); break;
        case 12: predict_n<12>(out, res, coef, order, shift, blocksize); break;
        default: predict_n<0>(out, res, coef, order, shift, blocksize); break;
    }
}

int read_subframe(BitReader& br, int64_t* out, int blocksize, int bps) {
    if (br.bits(1)) return FLAC_E_RESERVED;
    const int type = (int)br.bits(6);
    int wasted = 0;
    if (br.bits(1)) wasted = (int)br.unary() + 1;
    if (br.fail) return FLAC_E_TRUNC;
    bps -= wasted;
    if (bps < 1) return FLAC_E_HEADER;
    if (type == 0) {                                        // CONSTANT
        const int64_t v = bps > 32 ? (int64_t)(((uint64_t)br.bits(bps - 32) << 32) | br.bits(32)) : (int64_t)br.sbits(bps);
        for (int i = 0; i < blocksize; ++i) out[i] = v;
    } else if (type == 1) {                                 // VERBATIM
        for (int i = 0; i < blocksize; ++i) out[i] = bps > 32 ? 0 : (int64_t)br.sbits(bps);
        if (bps > 32) return FLAC_E_UNSUPPORTED;
    } else if ((type >= 8 && type <= 12) || type >= 32) {
        const bool lpc = type >= 32;
        const int order = lpc ? (type & 31) + 1 : type - 8;
        if (order > blocksize) return FLAC_E_HEADER;
        if (bps > 32) {                                     // 33-bit side channel of a 32-bit stream: warm-up needs 33 bits
            for (int i = 0; i < order; ++i) { const int64_t hi = (int64_t)br.sbits(bps - 32); out[i] = (hi << 32) | br.bits(32); }
        } else {
            for (int i = 0; i < order; ++i) out[i] = br.sbits(bps);
        }
        int32_t coef[32]; int shift = 0;
        if (lpc) {
            const int prec = (int)br.bits(4) + 1;
            if (prec == 16) return FLAC_E_RESERVED;
            shift = br.sbits(5);
            if (shift < 0) return FLAC_E_RESERVED;
            for (int i = 0; i < order; ++i) coef[i] = br.sbits(prec);
        } else {
            static const int32_t fixed[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
            for (int i = 0; i < order; ++i) coef[i] = fixed[order][i];
        }
        std::vector<int32_t> res((size_t)blocksize);
        const int rc = read_residual(br, res.data() + order, blocksize, order);
        if (rc != FLAC_OK) return rc;
        predict(out, res.data(), coef, order, shift, blocksize);
    } else {
        return FLAC_E_RESERVED;
    }
    if (br.fail) return FLAC_E_TRUNC;
    if (wasted) for (int i = 0; i < blocksize; ++i) out[i] = (int64_t)((uint64_t)out[i] << wasted);
    return FLAC_OK;
}

struct Decoded {
    StreamInfo si; std::vector<int32_t> pcm; int64_t samples = 0; bool md5_checked = false;   // pcm interleaved
    int16_t* mono16 = nullptr; int64_t mono16_cap = 0;    // optional direct sink: channel mean as int16 (16-bit streams), `pcm` stays empty
};

int decode(const uint8_t* data, int64_t n, int64_t max_samples, bool verify_md5, Decoded& d, bool info_only = false) {
    init_tables();
    if (n >= 10 && memcmp(data, "ID3", 3) == 0) {           // an ID3v2 tag some taggers put in front of the stream: 10-byte header, sync-safe size
        const int64_t sz = ((int64_t)(data[6] & 0x7f) << 21) | ((int64_t)(data[7] & 0x7f) << 14) | ((int64_t)(data[8] & 0x7f) << 7) | (data[9] & 0x7f);
        const int64_t skip = 10 + sz + ((data[5] & 0x10) ? 10 : 0);
        if (skip >= n) return FLAC_E_TRUNC;
        data += skip; n -= skip;
    }
    if (n < 42 || memcmp(data, "fLaC", 4) != 0) return FLAC_E_MAGIC;
    int64_t pos = 4; bool last = false, have_si = false;
    while (!last) {
        if (pos + 4 > n) return FLAC_E_TRUNC;
        last = (data[pos] & 0x80) != 0;
        const int type = data[pos] & 0x7f;
        const int64_t len = ((int64_t)data[pos + 1] << 16) | ((int64_t)data[pos + 2] << 8) | data[pos + 3];
        pos += 4;
        if (pos + len > n) return FLAC_E_TRUNC;
        if (type == 127) return FLAC_E_RESERVED;
        if (type == 0) {
            if (len != 34 || have_si) return FLAC_E_STREAMINFO;
            const uint8_t* s = data + pos;
            d.si.min_block = (s[0] << 8) | s[1]; d.si.max_block = (s[2] << 8) | s[3];
            d.si.rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            d.si.channels = ((s[12] >> 1) & 7) + 1;
            d.si.bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            d.si.total = ((int64_t)(s[13] & 15) << 32) | ((int64_t)s[14] << 24) | ((int64_t)s[15] << 16) | ((int64_t)s[16] << 8) | s[17];
            memcpy(d.si.md5, s + 18, 16);
            have_si = true;
        } else if (!have_si) {
            return FLAC_E_STREAMINFO;                       // STREAMINFO must be the first metadata block
        }
        pos += len;
    }
    if (!have_si || d.si.bps < 4 || d.si.bps > 32 || d.si.max_block < 16 || d.si.min_block > d.si.max_block) return FLAC_E_STREAMINFO;
    const int ch = d.si.channels;
    const int64_t want = (max_samples > 0 && (d.si.total == 0 || max_samples < d.si.total)) ? max_samples : (d.si.total > 0 ? d.si.total : INT64_MAX);
    const bool whole = d.si.total == 0 || want >= d.si.total;
    const size_t stride = (size_t)(d.si.max_block > 16 ? d.si.max_block : 16);
    std::vector<int64_t> chan((size_t)ch * stride);
    if (info_only) return FLAC_OK;
    if (want != INT64_MAX && !d.mono16) d.pcm.reserve((size_t)(want < (1 << 24) ? want : (1 << 24)) * ch);   // STREAMINFO is untrusted input
    Md5 md5;
    std::vector<uint8_t> md5_buf;
    bool cut_short = false;      // a stream of unknown length (total == 0) stopped by max_samples: the MD5 covers only a head
    d.pcm.clear();
    d.samples = 0;
    while (pos < n && d.samples < want) {
        // ---------------- frame header ----------------
        if (pos + 5 > n) return FLAC_E_TRUNC;
        if (data[pos] != 0xFF || (data[pos + 1] & 0xFE) != 0xF8) {
            if (d.si.total == 0 && d.samples > 0) break;     // stream of unknown length followed by something else (e.g. an ID3v1 tag)
            return FLAC_E_HEADER;
        }
        BitReader br(data + pos, n - pos);
        br.bits(15);
        br.bits(1);                                          // blocking strategy: the sample position is not needed for a sequential decode
        const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4), ch_code = (int)br.bits(4), ss_code = (int)br.bits(3);
        if (br.bits(1)) return FLAC_E_RESERVED;
        if (bs_code == 0 || sr_code == 15 || ch_code > 10 || ss_code == 3) return FLAC_E_RESERVED;
        {   // UTF-8-like coded frame / sample number: 1..7 bytes
            const uint32_t b0 = br.bits(8);
            int extra = 0;
            if (b0 == 0xFF) return FLAC_E_HEADER;
            if (b0 & 0x80) { uint32_t m = 0x40; while (b0 & m) { ++extra; m >>= 1; } if (extra == 0) return FLAC_E_HEADER; }
            for (int i = 0; i < extra; ++i) if ((br.bits(8) & 0xC0) != 0x80) return FLAC_E_HEADER;
        }
        int blocksize;
        if (bs_code == 1) blocksize = 192;
        else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
        else if (bs_code == 6) blocksize = (int)br.bits(8) + 1;
        else if (bs_code == 7) blocksize = (int)br.bits(16) + 1;
        else blocksize = 256 << (bs_code - 8);
        if (sr_code == 12) br.bits(8); else if (sr_code == 13 || sr_code == 14) br.bits(16);
        if (br.fail) return FLAC_E_TRUNC;
        const int64_t hdr_bytes = br.bitpos() >> 3;
        if (crc8(data + pos, hdr_bytes) != br.bits(8)) return FLAC_E_CRC8;
        static const int ss_table[8] = {0, 8, 12, 0, 16, 20, 24, 32};
        const int bps = ss_code == 0 ? d.si.bps : ss_table[ss_code];
        const int nch = ch_code < 8 ? ch_code + 1 : 2;
        if (nch != ch || bps != d.si.bps || (size_t)blocksize > stride) return FLAC_E_UNSUPPORTED;
        if (bps == 32 && ch_code >= 8) return FLAC_E_UNSUPPORTED;   // 33-bit side channel: not needed for speech corpora, rejected rather than guessed     // mid-stream format changes: not in scope
        // ---------------- subframes ----------------
        for (int c = 0; c < nch; ++c) {
            const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
            const int rc = read_subframe(br, chan.data() + (size_t)c * stride, blocksize, bps + (side ? 1 : 0));
            if (rc != FLAC_OK) return rc;
        }
        br.align();
        const int64_t body = br.bitpos() >> 3;
        if (pos + body + 2 > n) return FLAC_E_TRUNC;
        if (crc16(data + pos, body) != (uint16_t)((data[pos + body] << 8) | data[pos + body + 1])) return FLAC_E_CRC16;
        pos += body + 2;
        int64_t* L = chan.data(); int64_t* R = chan.data() + stride;
        if (ch_code == 8) for (int i = 0; i < blocksize; ++i) R[i] = (int64_t)((uint64_t)L[i] - (uint64_t)R[i]);          // left / side
        else if (ch_code == 9) for (int i = 0; i < blocksize; ++i) L[i] = (int64_t)((uint64_t)L[i] + (uint64_t)R[i]);     // side / right
        else if (ch_code == 10) for (int i = 0; i < blocksize; ++i) {                                   // mid / side
            const uint64_t s = (uint64_t)R[i], m = ((uint64_t)L[i] << 1) | (s & 1);
            L[i] = (int64_t)(m + s) >> 1; R[i] = (int64_t)(m - s) >> 1;
        }
        // ---------------- output (+ MD5 over the whole stream when it is decoded completely) ----------------
        const int64_t take = blocksize < want - d.samples ? blocksize : want - d.samples;
        if (take < blocksize && d.si.total == 0) cut_short = true;
        if (d.mono16) {
            if (bps != 16) return FLAC_E_UNSUPPORTED;
            if (d.samples + take > d.mono16_cap) return FLAC_E_ARG;
            int16_t* o = d.mono16 + d.samples;
            if (ch == 1) {
                for (int64_t i = 0; i < take; ++i) o[i] = (int16_t)chan[(size_t)i];
            } else {
                for (int64_t i = 0; i < take; ++i) {
                    int64_t sum = 0;
                    for (int c = 0; c < ch; ++c) sum += chan[(size_t)c * stride + i];
                    const int64_t a = sum < 0 ? -sum : sum, r = (2 * a + ch) / (2 * ch);     // mean over channels, halves away from zero (as ingest.read_wav_pcm16)
                    o[i] = (int16_t)(sum < 0 ? -r : r);
                }
            }
        } else {
            const size_t base = d.pcm.size();
            d.pcm.resize(base + (size_t)take * ch);
            for (int64_t i = 0; i < take; ++i)
                for (int c = 0; c < ch; ++c) d.pcm[base + (size_t)i * ch + c] = (int32_t)chan[(size_t)c * stride + i];
        }
        if (whole && verify_md5) {
            const int bytes = (bps + 7) / 8;
            md5_buf.resize((size_t)take * ch * bytes);
            uint8_t* t = md5_buf.data();
            if (bytes == 2 && ch == 1) {
                for (int64_t i = 0; i < take; ++i) { const uint32_t v = (uint32_t)chan[(size_t)i]; t[2 * i] = (uint8_t)v; t[2 * i + 1] = (uint8_t)(v >> 8); }
            } else {
                size_t k = 0;
                for (int64_t i = 0; i < take; ++i)
                    for (int c = 0; c < ch; ++c) { const uint32_t v = (uint32_t)chan[(size_t)c * stride + i]; for (int b = 0; b < bytes; ++b) t[k++] = (uint8_t)(v >> (8 * b)); }
            }
            md5.update(md5_buf.data(), md5_buf.size());
        }
        d.samples += take;
    }
    if (d.si.total > 0 && whole && d.samples != d.si.total) return FLAC_E_TRUNC;
    if (d.si.total == 0 && d.samples == 0) return FLAC_E_TRUNC;
    // total == 0 with max_samples: `whole` could not know the length; if the loop stopped at `want` with audio frames still
    // ahead (or inside a block), the digest is of a truncated head and must not be compared with STREAMINFO's
    if (d.si.total == 0 && max_samples > 0 && d.samples >= want && pos + 1 < n && data[pos] == 0xFF && (data[pos + 1] & 0xFE) == 0xF8) cut_short = true;
    if (whole && verify_md5 && !cut_short) {
        bool zero = true;
        for (int i = 0; i < 16; ++i) zero = zero && d.si.md5[i] == 0;
        if (!zero) {                                         // an all-zero signature means "not computed by the encoder"
            uint8_t got[16];
            md5.final(got);
            if (memcmp(got, d.si.md5, 16) != 0) return FLAC_E_MD5;
            d.md5_checked = true;
        }
    }
    return FLAC_OK;
}

// ---- frame table for the device decoder (flac_gpu.cu) --------------------------------------------------------------
// Parses a frame header at data[pos..]; returns its length in bytes (incl. the CRC-8) or 0 if this is not a valid header.
// number: the coded frame number (fixed block size) or first sample number (variable block size).
int64_t parse_frame_header(const uint8_t* data, int64_t n, int64_t pos, int* blocksize, int* variable, uint64_t* number, int* ch_code, int* ss_code) {
    if (pos + 5 > n || data[pos] != 0xFF || (data[pos + 1] & 0xFE) != 0xF8) return 0;
    BitReader br(data + pos, n - pos);
    br.bits(15);
    *variable = (int)br.bits(1);
    const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4);
    *ch_code = (int)br.bits(4); *ss_code = (int)br.bits(3);
    if (br.bits(1)) return 0;
    if (bs_code == 0 || sr_code == 15 || *ch_code > 10 || *ss_code == 3) return 0;
    const uint32_t b0 = br.bits(8);
    int extra = 0;
    if (b0 == 0xFF) return 0;
    uint64_t v = b0;
    if (b0 & 0x80) {
        uint32_t m = 0x40;
        while (b0 & m) { ++extra; m >>= 1; }
        if (extra == 0) return 0;
        v = b0 & (m - 1);
    }
    for (int i = 0; i < extra; ++i) { const uint32_t c = br.bits(8); if ((c & 0xC0) != 0x80) return 0; v = (v << 6) | (c & 0x3F); }
    *number = v;
    if (bs_code == 1) *blocksize = 192;
    else if (bs_code <= 5) *blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) *blocksize = (int)br.bits(8) + 1;
    else if (bs_code == 7) *blocksize = (int)br.bits(16) + 1;
    else *blocksize = 256 << (bs_code - 8);
    if (sr_code == 12) br.bits(8); else if (sr_code == 13 || sr_code == 14) br.bits(16);
    if (br.fail) return 0;
    const int64_t hdr = br.bitpos() >> 3;
    if (pos + hdr + 1 > n || crc8(data + pos, hdr) != data[pos + hdr]) return 0;
    return hdr + 1;
}

}  // namespace

extern "C" {

int64_t slsb_flac_scan(const uint8_t* data, int64_t nbytes, int64_t max_samples, int32_t* info, int64_t* frame_off, int32_t* frame_len,
                       int32_t* frame_samples, int64_t cap) {
    if (!data || nbytes <= 0 || !info || !frame_off || !frame_len || !frame_samples || cap <= 0) return FLAC_E_ARG;
    init_tables();
    Decoded d;
    const int rc = decode(data, nbytes, 0, false, d, true);              // STREAMINFO (+ leading ID3v2 tag / metadata checks)
    info[0] = d.si.rate; info[1] = d.si.channels; info[2] = d.si.bps; info[3] = (int32_t)(d.si.total & 0x7fffffff);
    info[4] = d.si.max_block; info[5] = (int32_t)(d.si.total >> 31);
    if (rc != FLAC_OK) return rc;
    // first audio frame: behind the metadata blocks
    int64_t pos = 0;
    if (nbytes >= 10 && memcmp(data, "ID3", 3) == 0) {
        const int64_t sz = ((int64_t)(data[6] & 0x7f) << 21) | ((int64_t)(data[7] & 0x7f) << 14) | ((int64_t)(data[8] & 0x7f) << 7) | (data[9] & 0x7f);
        pos = 10 + sz + ((data[5] & 0x10) ? 10 : 0);
    }
    pos += 4;
    for (bool last = false; !last;) {
        last = (data[pos] & 0x80) != 0;
        pos += 4 + (((int64_t)data[pos + 1] << 16) | ((int64_t)data[pos + 2] << 8) | data[pos + 3]);
    }
    int64_t nf = 0, samples = 0;
    int bs = 0, var = 0, chc = 0, ssc = 0; uint64_t num = 0;
    int64_t hdr = parse_frame_header(data, nbytes, pos, &bs, &var, &num, &chc, &ssc);
    if (!hdr) return FLAC_E_HEADER;
    const int64_t want = max_samples > 0 ? max_samples : INT64_MAX;
    while (samples < want) {
        if (nf >= cap) return FLAC_E_ARG;
        if ((size_t)bs > (size_t)(d.si.max_block > 16 ? d.si.max_block : 16)) return FLAC_E_UNSUPPORTED;
        // the frame ends where the next frame's header starts (sync code, valid CRC-8, consecutive number) AND the two bytes in
        // front of it are the CRC-16 of everything since this frame's sync code; a sync-like pattern inside the frame fails both
        const uint64_t next_num = var ? num + (uint64_t)bs : num + 1;
        int64_t end = -1;
        int nbs = 0, nvar = 0, nch = 0, nss = 0; uint64_t nnum = 0; int64_t nhdr = 0;
        uint16_t crc = 0; int64_t crc_upto = pos;                          // running CRC-16 over data[pos .. crc_upto)
        for (int64_t q = pos + hdr + 2; q + 1 < nbytes; ++q) {
            const uint8_t* f = static_cast<const uint8_t*>(memchr(data + q, 0xFF, (size_t)(nbytes - 1 - q)));
            if (!f) break;
            q = f - data;
            if ((data[q + 1] & 0xFE) != 0xF8) continue;
            nhdr = parse_frame_header(data, nbytes, q, &nbs, &nvar, &nnum, &nch, &nss);
            if (!nhdr || nvar != var || nnum != next_num) continue;
            if (crc_upto < q - 2) { crc = crc16_update(crc, data + crc_upto, q - 2 - crc_upto); crc_upto = q - 2; }
            if (crc_upto == q - 2 && crc == (uint16_t)((data[q - 2] << 8) | data[q - 1])) { end = q; break; }
        }
        bool is_last = false;
        if (end < 0) {                                                      // last frame of the stream: it ends with the data
            if (crc16(data + pos, nbytes - pos - 2) != (uint16_t)((data[nbytes - 2] << 8) | data[nbytes - 1])) return FLAC_E_CRC16;
            end = nbytes; is_last = true;
        }
        frame_off[nf] = pos; frame_len[nf] = (int32_t)(end - pos); frame_samples[nf] = bs;
        if (chc != 0) info[1] = info[1] > 1 ? info[1] : 2;                 // a multi-channel frame: the device decoder is mono only
        ++nf; samples += bs;
        if (is_last) break;
        pos = end; hdr = nhdr; bs = nbs; num = nnum; chc = nch; ssc = nss;
    }
    info[6] = (int32_t)(samples < want ? samples : want);
    (void)ssc;
    return nf;
}

int64_t slsb_flac_decode(const uint8_t* data, int64_t nbytes, int64_t max_samples, int verify_md5, int32_t* pcm_out, int64_t pcm_capacity,
                         int32_t* info) {
    if (!data || nbytes <= 0 || !info) return FLAC_E_ARG;
    Decoded d;
    int rc;
    try { rc = decode(data, nbytes, max_samples, verify_md5 != 0, d, pcm_out == nullptr); } catch (const std::bad_alloc&) { rc = FLAC_E_ARG; }
    info[0] = d.si.rate; info[1] = d.si.channels; info[2] = d.si.bps; info[3] = (int32_t)(d.si.total & 0x7fffffff);
    info[4] = d.md5_checked ? 1 : 0; info[5] = (int32_t)(d.si.total >> 31);
    if (rc != FLAC_OK) return rc;
    if (pcm_out) {
        if ((int64_t)d.pcm.size() > pcm_capacity) return FLAC_E_ARG;
        memcpy(pcm_out, d.pcm.data(), d.pcm.size() * sizeof(int32_t));
    }
    return d.samples;
}

int64_t slsb_flac_decode_mono16(const uint8_t* data, int64_t nbytes, int64_t max_samples, int verify_md5, int16_t* pcm_out,
                                int64_t pcm_capacity, int32_t* sample_rate) {
    if (!data || nbytes <= 0 || !pcm_out || pcm_capacity < 0) return FLAC_E_ARG;
    Decoded d;
    d.mono16 = pcm_out; d.mono16_cap = pcm_capacity;
    int rc;
    try { rc = decode(data, nbytes, max_samples, verify_md5 != 0, d); } catch (const std::bad_alloc&) { rc = FLAC_E_ARG; }
    if (sample_rate) *sample_rate = d.si.rate;
    if (rc != FLAC_OK) return rc;
    return d.samples;
}

}  // extern "C"
