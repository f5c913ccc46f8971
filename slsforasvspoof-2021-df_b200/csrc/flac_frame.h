// One FLAC frame -> 16-bit mono PCM: the frame-level core of the device decoder (flac_gpu.cu), written so that the SAME code
// compiles for the GPU (one thread per frame) and for the host (the CPU twin the `-m "not gpu"` tests pin against
// flac_decode.cpp and the RFC 9639 example files).  Format: RFC 9639 - frame header (section 9.1), subframes CONSTANT /
// VERBATIM / FIXED / LPC (9.2), partitioned Rice residual incl. escape partitions (9.2.7), wasted bits.  Scope: what the audio
// ingest needs (data_utils_SSL.py:109-113 decodes one mono 16-bit FLAC per item): 1 channel, 16 bits
// per sample; everything else is reported (FLACF_UNSUPPORTED) and the caller falls back to the host decoder.
//
// Integrity: frame boundaries, header CRC-8 and frame CRC-16 are established by the host scan (slsb_flac_scan, flac_decode.cpp)
// that produces the frame table - it reads every byte once at memory speed; the arithmetic-heavy part (Rice decode + predictor
// restore, ~50 operations per sample) runs here.  The core still checks what it can see itself: reserved bit patterns, bit-stream
// overrun, and that the frame ends exactly where the table says it does.
//
// Decoding is streamed: a sample's residual is Rice-decoded, the predictor (coefficients zero-padded to a fixed order, history
// in a register shift chain) is applied, and the int16 is stored - no residual buffer, no per-frame scratch memory.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FLACF_HD __host__ __device__ __forceinline__
#else
#define FLACF_HD inline
#endif

namespace flacf {

enum { FLACF_OK = 0, FLACF_TRUNC = -3, FLACF_HEADER = -4, FLACF_RESERVED = -7, FLACF_UNSUPPORTED = -9, FLACF_LENGTH = -12, FLACF_RANGE = -13 };

// MSB-first bit reader over 32-bit words of an arbitrarily aligned byte range; the buffer must stay readable up to the next
// 4-byte boundary after the range (the callers pad their buffers by 8 bytes).
struct Bits {
    const uint32_t* w;       // next aligned word to load
    uint64_t win;            // `cnt` valid bits at the top
    int cnt;
    int64_t left;            // bits of the range not yet loaded into the window
    bool fail;

    FLACF_HD static uint32_t load_be(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
        return __byte_perm(__ldg(p), 0, 0x0123);
#else
        const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
        return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
#endif
    }
    FLACF_HD void init(const uint8_t* p, int64_t nbytes) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const int skip = (int)(a & 3u);
        w = reinterpret_cast<const uint32_t*>(a - skip);
        win = 0; cnt = 0; fail = false;
        left = nbytes * 8;
        if (skip) {                                   // drop the bytes in front of the range
            const uint32_t v = load_be(w++);
            const int have = 32 - 8 * skip;
            win = (uint64_t)v << (32 + 8 * skip);
            cnt = have;
            if ((int64_t)have > left) { cnt = (int)left; win &= cnt ? ~0ull << (64 - cnt) : 0ull; }
            left -= cnt;
        }
    }
    FLACF_HD void refill() {                          // keeps >= 33 valid bits while the range lasts
        if (cnt <= 32 && left > 0) {
            uint64_t v = load_be(w++);
            int take = 32;
            if (left < 32) { take = (int)left; v &= ~0ull << (32 - take); }
            win |= v << (32 - cnt);
            cnt += take; left -= take;
        }
    }
    FLACF_HD uint32_t get(int k) {                    // k in 0..32
        if (k == 0) return 0;
        refill();
        if (cnt < k) { fail = true; cnt = 0; win = 0; return 0; }
        const uint32_t v = (uint32_t)(win >> (64 - k));
        win <<= k; cnt -= k;
        return v;
    }
    FLACF_HD int32_t sget(int k) {                    // two's complement, k in 1..32
        const uint32_t v = get(k);
        if (k == 32) return (int32_t)v;
        const uint32_t m = 1u << (k - 1);
        return (int32_t)((v ^ m) - m);
    }
    FLACF_HD static int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
        return __clzll((long long)x);
#else
        return __builtin_clzll(x);
#endif
    }
    FLACF_HD uint32_t unary() {                       // zeros before the next one bit (consumed too)
        uint32_t z = 0;
        for (;;) {
            refill();
            if (cnt == 0) { fail = true; return 0; }
            if (win == 0) { z += (uint32_t)cnt; cnt = 0; continue; }     // bits below `cnt` are zero by construction
            const int lead = clz64(win);
            z += (uint32_t)lead;
            win = lead == 63 ? 0 : win << (lead + 1);
            cnt -= lead + 1;
            return z;
        }
    }
    FLACF_HD uint32_t rice(int k) {                   // one Rice code word with parameter k (0..30): quotient in unary, k low bits
        refill();
        const int lead = win ? clz64(win) : 64;
        if (lead + 1 + k <= cnt) {                    // whole code word inside the window (cnt <= 64, so lead + 1 + k < 65)
            const uint64_t s = lead == 63 ? 0 : win << (lead + 1);
            const uint32_t u = ((uint32_t)lead << k) | (k ? (uint32_t)(s >> (64 - k)) : 0u);
            win = k ? s << k : s; cnt -= lead + 1 + k;
            return u;
        }
        const uint32_t q = unary();
        return (q << k) | (k ? get(k) : 0u);
    }
    FLACF_HD int64_t consumed(int64_t nbytes) const { return nbytes * 8 - left - cnt; }     // bits read so far
    FLACF_HD void align() { const int r = cnt & 7; win <<= r; cnt -= r; }                    // `left` is a multiple of 8
};

// Sink of decoded samples: applies the wasted-bits shift and stores the first `keep` as int16.
struct Out16 {
    int16_t* out; int keep; int wasted; int i;
    FLACF_HD void put(int32_t v) {
        if (i < keep) out[i] = (int16_t)(int32_t)((uint32_t)v << wasted);
        ++i;
    }
};

// Residual partitions + predictor, streamed: ONE loop over the samples of the block (partition headers are read when the running
// count reaches zero), so the 32 frames a warp decodes walk the same trip count whatever their partition orders are.
// `hist[0]` is the most recent sample; coefficients beyond `order` are zero, so TAPS only has to be >= order.
// 32-bit datapath: samples, coefficients and the prediction sum are int32 - what libFLAC itself does when
// bps + precision + log2(order) <= 32 (every 16-bit stream an encoder produces).  `maxabs` tracks the largest |sample| so the caller
// can PROVE afterwards that no 32-bit sum wrapped (sum |coef| * maxabs < 2^31); a stream where it might have is handed to the
// host decoder (64-bit sums), so the device never returns a sample that differs from flac_decode.cpp.
template <int TAPS>
FLACF_HD int decode_predicted(Bits& br, Out16& o, int blocksize, int order, const int32_t* coef_in, int shift, const int32_t* warm,
                              uint32_t& maxabs) {
    int32_t hist[TAPS], coef[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) { hist[j] = 0; coef[j] = j < order ? coef_in[j] : 0; }
    for (int i = 0; i < order; ++i) {                 // warm-up samples, oldest first
#pragma unroll
        for (int j = TAPS - 1; j > 0; --j) hist[j] = hist[j - 1];
        hist[0] = warm[i];
        const uint32_t a = (uint32_t)(warm[i] < 0 ? -(int64_t)warm[i] : warm[i]);
        maxabs = a > maxabs ? a : maxabs;
        o.put(warm[i]);
    }
    const uint32_t method = br.get(2);
    if (method > 1) return FLACF_RESERVED;
    const int pbits = method == 0 ? 4 : 5, esc = method == 0 ? 15 : 31;
    const int porder = (int)br.get(4);
    const int parts = 1 << porder;
    if (porder > 0 && ((blocksize >> porder) << porder) != blocksize) return FLACF_HEADER;
    if ((blocksize >> porder) < order) return FLACF_HEADER;
    int part = 0, remaining = 0, k = 0, raw = 0;
    bool escape = false;
    for (int i = order; i < blocksize; ++i) {
        while (remaining == 0) {                       // next partition header (a partition may be empty: order == its length)
            if (part == parts) return FLACF_HEADER;
            remaining = (blocksize >> porder) - (part == 0 ? order : 0);
            k = (int)br.get(pbits);
            escape = k == esc;
            raw = escape ? (int)br.get(5) : 0;
            ++part;
        }
        --remaining;
        int32_t r;
        if (escape) {
            r = raw ? br.sget(raw) : 0;
        } else {
            const uint32_t u = br.rice(k);
            r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);                 // zig-zag
        }
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;                       // four partial sums (wrapping): shorter dependency chains
#pragma unroll
        for (int j = 0; j < TAPS; j += 4) {
            a0 += (uint32_t)coef[j] * (uint32_t)hist[j];
            if (j + 1 < TAPS) a1 += (uint32_t)coef[j + 1] * (uint32_t)hist[j + 1];
            if (j + 2 < TAPS) a2 += (uint32_t)coef[j + 2] * (uint32_t)hist[j + 2];
            if (j + 3 < TAPS) a3 += (uint32_t)coef[j + 3] * (uint32_t)hist[j + 3];
        }
        const int32_t sum = (int32_t)((a0 + a1) + (a2 + a3));
        const int32_t v = (int32_t)((uint32_t)(sum >> shift) + (uint32_t)r);
#pragma unroll
        for (int j = TAPS - 1; j > 0; --j) hist[j] = hist[j - 1];
        hist[0] = v;
        const uint32_t a = (uint32_t)(v < 0 ? -(int64_t)v : v);
        maxabs = a > maxabs ? a : maxabs;
        o.put(v);
    }
    if (br.fail) return FLACF_TRUNC;
    while (part < parts) {                             // trailing empty partitions still carry their headers
        if ((blocksize >> porder) - (part == 0 ? order : 0) != 0) return FLACF_HEADER;
        const int kk = (int)br.get(pbits);
        if (kk == esc) br.get(5);
        ++part;
    }
    return br.fail ? FLACF_TRUNC : FLACF_OK;
}

// One mono subframe of `bps` (<= 17) bits -> o
FLACF_HD int decode_subframe(Bits& br, Out16& o, int blocksize, int bps) {
    if (br.get(1)) return FLACF_RESERVED;
    const int type = (int)br.get(6);
    int wasted = 0;
    if (br.get(1)) wasted = (int)br.unary() + 1;
    if (br.fail) return FLACF_TRUNC;
    bps -= wasted;
    if (bps < 1) return FLACF_HEADER;
    o.wasted = wasted;
    if (type == 0) {                                        // CONSTANT
        const int32_t v = br.sget(bps);
        for (int i = 0; i < blocksize; ++i) o.put(v);
    } else if (type == 1) {                                 // VERBATIM
        for (int i = 0; i < blocksize; ++i) o.put(br.sget(bps));
    } else if ((type >= 8 && type <= 12) || type >= 32) {
        const bool lpc = type >= 32;
        const int order = lpc ? (type & 31) + 1 : type - 8;
        if (order > blocksize) return FLACF_HEADER;
        int32_t warm[32];
        for (int i = 0; i < order; ++i) warm[i] = br.sget(bps);
        int32_t coef[32];
        int shift = 0;
        if (lpc) {
            const int prec = (int)br.get(4) + 1;
            if (prec == 16) return FLACF_RESERVED;
            shift = br.sget(5);
            if (shift < 0) return FLACF_RESERVED;
            for (int i = 0; i < order; ++i) coef[i] = br.sget(prec);
        } else {
            const int32_t fixed[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
            for (int i = 0; i < order; ++i) coef[i] = fixed[order][i];
        }
        if (br.fail) return FLACF_TRUNC;
        uint64_t sumabs = 0;
        for (int i = 0; i < order; ++i) sumabs += (uint64_t)(coef[i] < 0 ? -(int64_t)coef[i] : coef[i]);
        uint32_t maxabs = 0;
        int rc;
        if (order <= 4) rc = decode_predicted<4>(br, o, blocksize, order, coef, shift, warm, maxabs);
        else if (order <= 8) rc = decode_predicted<8>(br, o, blocksize, order, coef, shift, warm, maxabs);
        else if (order <= 12) rc = decode_predicted<12>(br, o, blocksize, order, coef, shift, warm, maxabs);
        else rc = decode_predicted<32>(br, o, blocksize, order, coef, shift, warm, maxabs);
        if (rc != FLACF_OK) return rc;
        // proof that the 32-bit sums never wrapped and every sample stayed an int32: otherwise the host decoder takes the clip
        if (sumabs * (uint64_t)maxabs >= (1ull << 31) || maxabs >= (1u << 30)) return FLACF_RANGE;
    } else {
        return FLACF_RESERVED;
    }
    return br.fail ? FLACF_TRUNC : FLACF_OK;
}

// A whole frame: header (already CRC-checked by the scan), one subframe, padding, CRC-16 position check.
// p / nbytes: the frame from its sync code up to and including its CRC-16; bps_si: STREAMINFO sample size.
// Writes min(blocksize, keep) samples to out; returns the frame's block size (> 0) or an error (< 0).
FLACF_HD int decode_frame_mono16(const uint8_t* p, int64_t nbytes, int bps_si, int keep, int16_t* out) {
    if (nbytes < 6) return FLACF_TRUNC;
    Bits br;
    br.init(p, nbytes);
    if (br.get(15) != 0x7FFC) return FLACF_HEADER;          // 14-bit sync code + the reserved zero bit
    br.get(1);                                              // blocking strategy
    const int bs_code = (int)br.get(4), sr_code = (int)br.get(4), ch_code = (int)br.get(4), ss_code = (int)br.get(3);
    if (br.get(1)) return FLACF_RESERVED;
    if (bs_code == 0 || sr_code == 15 || ch_code > 10 || ss_code == 3) return FLACF_RESERVED;
    {   // coded frame / sample number: 1..7 bytes
        const uint32_t b0 = br.get(8);
        int extra = 0;
        if (b0 == 0xFF) return FLACF_HEADER;
        if (b0 & 0x80) { uint32_t m = 0x40; while (b0 & m) { ++extra; m >>= 1; } if (extra == 0) return FLACF_HEADER; }
        for (int i = 0; i < extra; ++i) if ((br.get(8) & 0xC0) != 0x80) return FLACF_HEADER;
    }
    int blocksize;
    if (bs_code == 1) blocksize = 192;
    else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = (int)br.get(8) + 1;
    else if (bs_code == 7) blocksize = (int)br.get(16) + 1;
    else blocksize = 256 << (bs_code - 8);
    if (sr_code == 12) br.get(8); else if (sr_code == 13 || sr_code == 14) br.get(16);
    br.get(8);                                              // CRC-8 (verified by the scan)
    if (br.fail) return FLACF_TRUNC;
    const int ss_table[8] = {0, 8, 12, 0, 16, 20, 24, 32};
    const int bps = ss_code == 0 ? bps_si : ss_table[ss_code];
    if (ch_code != 0 || bps != 16) return FLACF_UNSUPPORTED;
    Out16 o{out, keep, 0, 0};
    const int rc = decode_subframe(br, o, blocksize, bps);
    if (rc != FLACF_OK) return rc;
    br.align();
    if (br.consumed(nbytes) != (nbytes - 2) * 8) return FLACF_LENGTH;      // the subframe must end right in front of the CRC-16
    return blocksize;
}

}  // namespace flacf
