/* FLAC stream decoder (host side, plain C99) — libwb_audio.so, declared in include/wb_audio.h.
 *
 * Takes the place of libsndfile behind `datasets`' Audio feature in the reference's scripts (run.py:266-267): FLAC bytes ->
 * PCM.  Written from the format description (xiph.org FLAC format / RFC 9639): stream marker, metadata blocks, frames with
 * a CRC-8 protected header, CONSTANT / VERBATIM / FIXED / LPC subframes with partitioned Rice residuals, inter-channel
 * decorrelation, CRC-16 frame footer, MD5 signature of the PCM in STREAMINFO.  Everything is integer arithmetic; a decode
 * that passes the per-frame CRCs and the stream MD5 is bit-exact by construction.
 */
#include "../../../include/wb_audio.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ error reporting */
static __thread char g_error[256];

WB_AUDIO_API const char* wb_audio_last_error(void) { return g_error; }
WB_AUDIO_API int wb_audio_version(void) { return 100; }

#define FAIL(code, ...)                                   \
    do {                                                  \
        snprintf(g_error, sizeof g_error, __VA_ARGS__);   \
        return (code);                                    \
    } while (0)

/* ------------------------------------------------------------------------------------------------ MD5 (RFC 1321) */
typedef struct {
    uint32_t h[4];
    uint64_t length;
    uint8_t block[64];
    size_t fill;
} Md5;

static const uint32_t kMd5K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
    0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
    0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
    0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
    0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
    0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t kMd5S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_init(Md5* m) {
    m->h[0] = 0x67452301u;
    m->h[1] = 0xefcdab89u;
    m->h[2] = 0x98badcfeu;
    m->h[3] = 0x10325476u;
    m->length = 0;
    m->fill = 0;
}

static void md5_block(Md5* m, const uint8_t* p) {
    uint32_t w[16];
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
    uint32_t a = m->h[0], b = m->h[1], c = m->h[2], d = m->h[3];
    for (int i = 0; i < 64; ++i) {
        uint32_t f;
        int g;
        if (i < 16) {
            f = (b & c) | (~b & d);
            g = i;
        } else if (i < 32) {
            f = (d & b) | (~d & c);
            g = (5 * i + 1) & 15;
        } else if (i < 48) {
            f = b ^ c ^ d;
            g = (3 * i + 5) & 15;
        } else {
            f = c ^ (b | ~d);
            g = (7 * i) & 15;
        }
        uint32_t t = a + f + kMd5K[i] + w[g];
        a = d;
        d = c;
        c = b;
        b += (t << kMd5S[i]) | (t >> (32 - kMd5S[i]));
    }
    m->h[0] += a;
    m->h[1] += b;
    m->h[2] += c;
    m->h[3] += d;
}

static void md5_update(Md5* m, const uint8_t* p, size_t n) {
    m->length += n;
    if (m->fill) {
        size_t take = 64 - m->fill < n ? 64 - m->fill : n;
        memcpy(m->block + m->fill, p, take);
        m->fill += take;
        p += take;
        n -= take;
        if (m->fill < 64) return;
        md5_block(m, m->block);
        m->fill = 0;
    }
    for (; n >= 64; p += 64, n -= 64) md5_block(m, p);
    if (n) {
        memcpy(m->block, p, n);
        m->fill = n;
    }
}

static void md5_final(Md5* m, uint8_t out[16]) {
    uint64_t bits = m->length * 8;
    uint8_t pad[72] = {0x80};
    size_t padlen = (m->fill < 56 ? 56 : 120) - m->fill;
    for (int i = 0; i < 8; ++i) pad[padlen + i] = (uint8_t)(bits >> (8 * i));
    md5_update(m, pad, padlen + 8);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(m->h[i] >> (8 * j));
}

WB_AUDIO_API int wb_md5(const uint8_t* data, size_t size, uint8_t digest[16]) {
    if ((!data && size) || !digest) FAIL(WB_AUDIO_ERR_ARGUMENT, "wb_md5: null pointer");
    Md5 m;
    md5_init(&m);
    md5_update(&m, data, size);
    md5_final(&m, digest);
    return WB_AUDIO_OK;
}

/* ------------------------------------------------------------------------------------------------ CRCs */
static uint8_t g_crc8[256];
static uint16_t g_crc16[256];
static int g_tables_ready;

static void build_tables(void) {
    if (g_tables_ready) return;
    for (int i = 0; i < 256; ++i) {
        uint8_t c8 = (uint8_t)i;                /* x^8 + x^2 + x + 1 */
        uint16_t c16 = (uint16_t)(i << 8);      /* x^16 + x^15 + x^2 + 1 */
        for (int b = 0; b < 8; ++b) {
            c8 = (uint8_t)((c8 & 0x80) ? (c8 << 1) ^ 0x07 : c8 << 1);
            c16 = (uint16_t)((c16 & 0x8000) ? (c16 << 1) ^ 0x8005 : c16 << 1);
        }
        g_crc8[i] = c8;
        g_crc16[i] = c16;
    }
    g_tables_ready = 1;   /* idempotent: a racing thread writes the same values */
}

static uint8_t crc8(const uint8_t* p, size_t n) {
    uint8_t c = 0;
    while (n--) c = g_crc8[c ^ *p++];
    return c;
}

static uint16_t crc16(const uint8_t* p, size_t n) {
    uint16_t c = 0;
    while (n--) c = (uint16_t)((c << 8) ^ g_crc16[(c >> 8) ^ *p++]);
    return c;
}

/* ------------------------------------------------------------------------------------------------ bit reader (MSB first) */
typedef struct {
    const uint8_t* start;
    const uint8_t* p;     /* next byte to load */
    const uint8_t* end;
    uint64_t acc;         /* unread bits, left aligned */
    int nbits;            /* number of valid bits in acc */
    int overrun;          /* bytes loaded from beyond `end` (read as zero; the reader looks up to 8 bytes ahead) */
} BitReader;

static void br_init(BitReader* b, const uint8_t* start, const uint8_t* end) {
    b->start = b->p = start;
    b->end = end;
    b->acc = 0;
    b->nbits = 0;
    b->overrun = 0;
}

static inline void br_refill(BitReader* b) {
    while (b->nbits <= 56) {
        uint64_t byte = 0;
        if (b->p < b->end)
            byte = *b->p;
        else
            b->overrun++;
        b->p++;
        b->acc |= byte << (56 - b->nbits);
        b->nbits += 8;
    }
}

static inline uint32_t br_read(BitReader* b, int n) { /* 0 <= n <= 32 */
    if (n == 0) return 0;
    br_refill(b);
    uint32_t v = (uint32_t)(b->acc >> (64 - n));
    b->acc <<= n;
    b->nbits -= n;
    return v;
}

static inline int64_t br_read_signed(BitReader* b, int n) { /* 0 <= n <= 33 */
    if (n == 0) return 0;
    uint64_t v;
    if (n > 32) {
        v = (uint64_t)br_read(b, n - 32) << 32;
        v |= br_read(b, 32);
    } else {
        v = br_read(b, n);
    }
    uint64_t sign = (uint64_t)1 << (n - 1);
    return (int64_t)((v ^ sign) - sign);
}

/* number of 0 bits before the next 1 bit (which is consumed); -1 when the data runs out */
static inline int64_t br_read_unary(BitReader* b) {
    int64_t zeros = 0;
    for (;;) {
        br_refill(b);
        if (b->acc == 0) {
            if (b->overrun > 8) return -1;
            zeros += b->nbits;
            b->nbits = 0;
            continue;
        }
        int z = __builtin_clzll(b->acc);
        zeros += z;
        b->acc <<= z;      /* z <= 63 */
        b->acc <<= 1;
        b->nbits -= z + 1;
        return zeros;
    }
}

static inline void br_align(BitReader* b) {
    int drop = b->nbits & 7;
    b->acc <<= drop;
    b->nbits -= drop;
}

/* byte offset (from `start`) of the next unread bit; call on a byte boundary */
static inline size_t br_byte_pos(const BitReader* b) { return (size_t)(b->p - b->start) - (size_t)(b->nbits >> 3); }

/* ------------------------------------------------------------------------------------------------ metadata */
static uint32_t be(const uint8_t* p, int n) {
    uint32_t v = 0;
    while (n--) v = v << 8 | *p++;
    return v;
}

WB_AUDIO_API int wb_flac_read_info(const uint8_t* data, size_t size, wb_flac_info* info) {
    if (!data || !info) FAIL(WB_AUDIO_ERR_ARGUMENT, "wb_flac_read_info: null pointer");
    size_t pos = 0;
    if (size >= 10 && memcmp(data, "ID3", 3) == 0) { /* ID3v2 tag: 10-byte header, 4 x 7-bit size, optional footer */
        size_t tag = ((size_t)(data[6] & 0x7f) << 21) | ((size_t)(data[7] & 0x7f) << 14) | ((size_t)(data[8] & 0x7f) << 7) | (data[9] & 0x7f);
        pos = 10 + tag + ((data[5] & 0x10) ? 10 : 0);
    }
    if (pos + 4 > size || memcmp(data + pos, "fLaC", 4) != 0) FAIL(WB_AUDIO_ERR_FORMAT, "not a FLAC stream (no fLaC marker)");
    pos += 4;
    int have_streaminfo = 0, last = 0, index = 0;
    while (!last) {
        if (pos + 4 > size) FAIL(WB_AUDIO_ERR_FORMAT, "truncated metadata block header");
        last = data[pos] >> 7;
        int type = data[pos] & 0x7f;
        size_t len = be(data + pos + 1, 3);
        pos += 4;
        if (pos + len > size) FAIL(WB_AUDIO_ERR_FORMAT, "truncated metadata block (type %d, %zu bytes)", type, len);
        if (type == 127) FAIL(WB_AUDIO_ERR_FORMAT, "invalid metadata block type 127");
        if (index == 0 && type != 0) FAIL(WB_AUDIO_ERR_FORMAT, "the first metadata block is not STREAMINFO");
        if (type == 0) {
            if (len != 34 || have_streaminfo) FAIL(WB_AUDIO_ERR_FORMAT, "bad STREAMINFO block");
            const uint8_t* s = data + pos;
            info->min_blocksize = be(s, 2);
            info->max_blocksize = be(s + 2, 2);
            uint64_t packed = ((uint64_t)be(s + 10, 4) << 32) | be(s + 14, 4); /* 20 rate | 3 ch-1 | 5 bps-1 | 36 total */
            info->sample_rate = (uint32_t)(packed >> 44);
            info->channels = (uint32_t)((packed >> 41) & 7) + 1;
            info->bits_per_sample = (uint32_t)((packed >> 36) & 31) + 1;
            info->total_samples = packed & (((uint64_t)1 << 36) - 1);
            memcpy(info->md5, s + 18, 16);
            have_streaminfo = 1;
        }
        pos += len;
        ++index;
    }
    if (!have_streaminfo) FAIL(WB_AUDIO_ERR_FORMAT, "no STREAMINFO block");
    if (info->bits_per_sample < 4) FAIL(WB_AUDIO_ERR_FORMAT, "invalid bits per sample %u", info->bits_per_sample);
    if (info->min_blocksize < 16 || info->max_blocksize < info->min_blocksize) FAIL(WB_AUDIO_ERR_FORMAT, "invalid block sizes in STREAMINFO");
    info->audio_offset = pos;
    return WB_AUDIO_OK;
}

/* ------------------------------------------------------------------------------------------------ frames */
typedef struct {
    uint32_t blocksize;
    uint32_t sample_rate;   /* 0 = as STREAMINFO */
    uint32_t bits;          /* 0 = as STREAMINFO */
    int assignment;         /* 0..7 independent channels (n + 1), 8 left/side, 9 side/right, 10 mid/side */
    uint64_t number;        /* frame number (fixed block size) or first sample number (variable) */
    int variable;
} FrameHeader;

static int read_frame_header(BitReader* b, FrameHeader* h) {
    size_t start = br_byte_pos(b);
    uint32_t sync = br_read(b, 15);
    if (sync != 0x7ffc) FAIL(WB_AUDIO_ERR_FORMAT, "lost frame sync at byte %zu", start);
    h->variable = (int)br_read(b, 1);
    int bs_code = (int)br_read(b, 4), sr_code = (int)br_read(b, 4);
    h->assignment = (int)br_read(b, 4);
    int size_code = (int)br_read(b, 3);
    if (br_read(b, 1)) FAIL(WB_AUDIO_ERR_FORMAT, "reserved bit set in the frame header at byte %zu", start);
    if (h->assignment > 10) FAIL(WB_AUDIO_ERR_FORMAT, "reserved channel assignment %d", h->assignment);
    /* frame / sample number, coded like UTF-8 extended to 36 bits */
    uint32_t first = br_read(b, 8);
    int extra;
    if (first < 0x80) {
        extra = 0;
        h->number = first;
    } else {
        int ones = 0;
        while (ones < 8 && (first & (0x80u >> ones))) ++ones;
        if (ones < 2 || ones > 7) FAIL(WB_AUDIO_ERR_FORMAT, "bad coded frame number");
        extra = ones - 1;
        h->number = first & (0xffu >> (ones + 1));
        if (ones == 7) h->number = 0;
    }
    for (int i = 0; i < extra; ++i) {
        uint32_t c = br_read(b, 8);
        if ((c & 0xc0) != 0x80) FAIL(WB_AUDIO_ERR_FORMAT, "bad continuation byte in the coded frame number");
        h->number = h->number << 6 | (c & 0x3f);
    }
    switch (bs_code) {
        case 0: FAIL(WB_AUDIO_ERR_FORMAT, "reserved block size code");
        case 1: h->blocksize = 192; break;
        case 2: case 3: case 4: case 5: h->blocksize = 576u << (bs_code - 2); break;
        case 6: h->blocksize = br_read(b, 8) + 1; break;
        case 7: h->blocksize = br_read(b, 16) + 1; break;
        default: h->blocksize = 256u << (bs_code - 8); break;
    }
    static const uint32_t kRates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    if (sr_code < 12) h->sample_rate = kRates[sr_code];
    else if (sr_code == 12) h->sample_rate = br_read(b, 8) * 1000;
    else if (sr_code == 13) h->sample_rate = br_read(b, 16);
    else if (sr_code == 14) h->sample_rate = br_read(b, 16) * 10;
    else FAIL(WB_AUDIO_ERR_FORMAT, "invalid sample rate code");
    static const uint32_t kBits[8] = {0, 8, 12, 0, 16, 20, 24, 32};
    if (size_code == 3) FAIL(WB_AUDIO_ERR_FORMAT, "reserved sample size code");
    h->bits = kBits[size_code];
    size_t crc_pos = br_byte_pos(b);
    uint32_t stored = br_read(b, 8);
    if (br_byte_pos(b) > (size_t)(b->end - b->start)) FAIL(WB_AUDIO_ERR_FORMAT, "truncated frame header at byte %zu", start);
    if (crc8(b->start + start, crc_pos - start) != stored) FAIL(WB_AUDIO_ERR_CHECKSUM, "frame header CRC-8 mismatch at byte %zu", start);
    return WB_AUDIO_OK;
}

static int read_residual(BitReader* b, int64_t* out, uint32_t blocksize, int order) {
    int method = (int)br_read(b, 2);
    if (method > 1) FAIL(WB_AUDIO_ERR_FORMAT, "reserved residual coding method %d", method);
    int param_bits = method ? 5 : 4, escape = method ? 31 : 15;
    int porder = (int)br_read(b, 4);
    uint32_t parts = 1u << porder;
    if ((blocksize & (parts - 1)) != 0 || (blocksize >> porder) < (uint32_t)order)
        FAIL(WB_AUDIO_ERR_FORMAT, "partition order %d does not fit block size %u / predictor order %d", porder, blocksize, order);
    uint32_t i = (uint32_t)order;
    for (uint32_t part = 0; part < parts; ++part) {
        uint32_t count = (blocksize >> porder) - (part == 0 ? (uint32_t)order : 0);
        int k = (int)br_read(b, param_bits);
        if (k == escape) {
            int raw = (int)br_read(b, 5);
            for (uint32_t j = 0; j < count; ++j) out[i++] = br_read_signed(b, raw);
        } else {
            for (uint32_t j = 0; j < count; ++j) {
                int64_t q = br_read_unary(b);
                if (q < 0) FAIL(WB_AUDIO_ERR_FORMAT, "data ends inside a Rice-coded residual");
                uint64_t u = ((uint64_t)q << k) | br_read(b, k);
                out[i++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
            }
        }
        if (b->overrun > 8) FAIL(WB_AUDIO_ERR_FORMAT, "data ends inside a residual partition");
    }
    return WB_AUDIO_OK;
}

static int read_subframe(BitReader* b, int64_t* s, uint32_t n, int bits) {
    if (br_read(b, 1)) FAIL(WB_AUDIO_ERR_FORMAT, "subframe padding bit set");
    int type = (int)br_read(b, 6);
    int wasted = 0;
    if (br_read(b, 1)) {
        int64_t z = br_read_unary(b);
        if (z < 0 || z + 1 >= bits) FAIL(WB_AUDIO_ERR_FORMAT, "bad wasted-bits count");
        wasted = (int)z + 1;
        bits -= wasted;
    }
    if (type == 0) { /* CONSTANT */
        int64_t v = br_read_signed(b, bits);
        for (uint32_t i = 0; i < n; ++i) s[i] = v;
    } else if (type == 1) { /* VERBATIM */
        for (uint32_t i = 0; i < n; ++i) s[i] = br_read_signed(b, bits);
    } else if (type >= 8 && type <= 12) { /* FIXED predictor, order 0..4 */
        int order = type - 8;
        if ((uint32_t)order > n) FAIL(WB_AUDIO_ERR_FORMAT, "predictor order exceeds the block size");
        for (int i = 0; i < order; ++i) s[i] = br_read_signed(b, bits);
        int rc = read_residual(b, s, n, order);
        if (rc) return rc;
        /* wrapping (unsigned) arithmetic: corrupt residuals must not trigger signed overflow before the CRC rejects the frame */
        uint64_t* u = (uint64_t*)s;
        switch (order) {
            case 1: for (uint32_t i = 1; i < n; ++i) u[i] += u[i - 1]; break;
            case 2: for (uint32_t i = 2; i < n; ++i) u[i] += 2 * u[i - 1] - u[i - 2]; break;
            case 3: for (uint32_t i = 3; i < n; ++i) u[i] += 3 * u[i - 1] - 3 * u[i - 2] + u[i - 3]; break;
            case 4: for (uint32_t i = 4; i < n; ++i) u[i] += 4 * u[i - 1] - 6 * u[i - 2] + 4 * u[i - 3] - u[i - 4]; break;
            default: break;
        }
    } else if (type >= 32) { /* LPC, order 1..32 */
        int order = type - 31;
        if ((uint32_t)order > n) FAIL(WB_AUDIO_ERR_FORMAT, "predictor order exceeds the block size");
        for (int i = 0; i < order; ++i) s[i] = br_read_signed(b, bits);
        int precision = (int)br_read(b, 4) + 1;
        if (precision == 16) FAIL(WB_AUDIO_ERR_FORMAT, "invalid LPC coefficient precision");
        int shift = (int)br_read_signed(b, 5);
        if (shift < 0) FAIL(WB_AUDIO_ERR_FORMAT, "negative LPC shift");
        int64_t coef[32];
        for (int i = 0; i < order; ++i) coef[i] = br_read_signed(b, precision);
        int rc = read_residual(b, s, n, order);
        if (rc) return rc;
        for (uint32_t i = (uint32_t)order; i < n; ++i) {
            uint64_t acc = 0; /* wrapping arithmetic, see above; exact for every valid stream (|sum| < 2^63) */
            for (int j = 0; j < order; ++j) acc += (uint64_t)coef[j] * (uint64_t)s[i - 1 - j];
            /* arithmetic shift: rounds towards minus infinity, as the format requires */
            s[i] = (int64_t)((uint64_t)s[i] + (uint64_t)((int64_t)acc >> shift));
        }
    } else {
        FAIL(WB_AUDIO_ERR_FORMAT, "reserved subframe type %d", type);
    }
    if (wasted)
        for (uint32_t i = 0; i < n; ++i) s[i] = (int64_t)((uint64_t)s[i] << wasted);
    return WB_AUDIO_OK;
}

WB_AUDIO_API int wb_flac_decode_i32(const uint8_t* data, size_t size, int32_t* out, uint64_t capacity, uint64_t* decoded, int verify_md5) {
    if (!data || !decoded) FAIL(WB_AUDIO_ERR_ARGUMENT, "wb_flac_decode_i32: null pointer");
    *decoded = 0;
    wb_flac_info info;
    int rc = wb_flac_read_info(data, size, &info);
    if (rc) return rc;
    build_tables();
    const uint32_t C = info.channels;
    int64_t* chan = (int64_t*)malloc((size_t)C * 65536 * sizeof(int64_t));
    uint8_t* pcm_bytes = NULL;
    const int bytes_per_sample = (int)(info.bits_per_sample + 7) / 8;
    int have_md5 = 0;
    for (int i = 0; i < 16; ++i) have_md5 |= info.md5[i];
    const int check_md5 = verify_md5 && have_md5;
    if (check_md5) pcm_bytes = (uint8_t*)malloc((size_t)C * 65536 * 4);
    if (!chan || (check_md5 && !pcm_bytes)) {
        free(chan);
        free(pcm_bytes);
        FAIL(WB_AUDIO_ERR_ARGUMENT, "out of memory");
    }
    Md5 md5;
    md5_init(&md5);
    BitReader b;
    br_init(&b, data, data + size);
    b.p = data + info.audio_offset;
    uint64_t done = 0;
    rc = WB_AUDIO_OK;
#define BAIL(code, ...)                                   \
    do {                                                  \
        snprintf(g_error, sizeof g_error, __VA_ARGS__);   \
        rc = (code);                                      \
        goto finish;                                      \
    } while (0)
    for (;;) {
        size_t frame_start = br_byte_pos(&b);
        if (frame_start >= size) break;
        if (info.total_samples && done >= info.total_samples) break;          /* trailing bytes (e.g. an ID3v1 tag) */
        if (size - frame_start < 2 || data[frame_start] != 0xff || (data[frame_start + 1] & 0xfe) != 0xf8) {
            if (done && !info.total_samples) break;                           /* unknown length: stop at the first non-frame */
            BAIL(WB_AUDIO_ERR_FORMAT, "lost frame sync at byte %zu", frame_start);
        }
        FrameHeader h;
        rc = read_frame_header(&b, &h);
        if (rc) goto finish;
        const uint32_t n = h.blocksize;
        const int bits = (int)(h.bits ? h.bits : info.bits_per_sample);
        const uint32_t frame_channels = h.assignment < 8 ? (uint32_t)h.assignment + 1 : 2;
        if (frame_channels != C || (uint32_t)bits != info.bits_per_sample)
            BAIL(WB_AUDIO_ERR_UNSUPPORTED, "frame at byte %zu changes the channel count or sample size mid-stream", frame_start);
        if (h.sample_rate && h.sample_rate != info.sample_rate)
            BAIL(WB_AUDIO_ERR_UNSUPPORTED, "frame at byte %zu changes the sample rate mid-stream", frame_start);
        if (n > info.max_blocksize && info.max_blocksize) BAIL(WB_AUDIO_ERR_FORMAT, "block of %u samples exceeds STREAMINFO's maximum %u", n, info.max_blocksize);
        for (uint32_t c = 0; c < C; ++c) {
            int side = (h.assignment == 8 && c == 1) || (h.assignment == 9 && c == 0) || (h.assignment == 10 && c == 1);
            rc = read_subframe(&b, chan + (size_t)c * 65536, n, bits + side);
            if (rc) goto finish;
        }
        br_align(&b);
        size_t crc_pos = br_byte_pos(&b);
        uint32_t stored = br_read(&b, 16);
        if (br_byte_pos(&b) > size) BAIL(WB_AUDIO_ERR_FORMAT, "data ends inside the frame that starts at byte %zu", frame_start);
        if (crc16(data + frame_start, crc_pos - frame_start) != stored) BAIL(WB_AUDIO_ERR_CHECKSUM, "frame CRC-16 mismatch (frame at byte %zu)", frame_start);
        int64_t* c0 = chan;
        int64_t* c1 = chan + 65536;
        if (h.assignment == 8) { /* left, side = left - right */
            for (uint32_t i = 0; i < n; ++i) c1[i] = (int64_t)((uint64_t)c0[i] - (uint64_t)c1[i]);
        } else if (h.assignment == 9) { /* side, right */
            for (uint32_t i = 0; i < n; ++i) c0[i] = (int64_t)((uint64_t)c0[i] + (uint64_t)c1[i]);
        } else if (h.assignment == 10) { /* mid = (left + right) >> 1, side = left - right: the dropped bit is side's parity */
            for (uint32_t i = 0; i < n; ++i) {
                uint64_t side = (uint64_t)c1[i], mid = ((uint64_t)c0[i] << 1) | (side & 1);
                c0[i] = (int64_t)(mid + side) >> 1;
                c1[i] = (int64_t)(mid - side) >> 1;
            }
        }
        if (info.total_samples && done + n > info.total_samples) BAIL(WB_AUDIO_ERR_FORMAT, "more samples than STREAMINFO announces");
        if (out) {
            if (done + n > capacity) BAIL(WB_AUDIO_ERR_ARGUMENT, "output buffer too small: %llu samples per channel, the stream has more", (unsigned long long)capacity);
            int32_t* dst = out + done * C;
            for (uint32_t i = 0; i < n; ++i)
                for (uint32_t c = 0; c < C; ++c) dst[(size_t)i * C + c] = (int32_t)chan[(size_t)c * 65536 + i];
        }
        if (check_md5) { /* the signature covers the interleaved samples, little endian, in whole bytes */
            uint8_t* q = pcm_bytes;
            for (uint32_t i = 0; i < n; ++i)
                for (uint32_t c = 0; c < C; ++c) {
                    uint64_t v = (uint64_t)chan[(size_t)c * 65536 + i];
                    for (int k = 0; k < bytes_per_sample; ++k) *q++ = (uint8_t)(v >> (8 * k));
                }
            md5_update(&md5, pcm_bytes, (size_t)(q - pcm_bytes));
        }
        done += n;
    }
    if (info.total_samples && done != info.total_samples)
        BAIL(WB_AUDIO_ERR_FORMAT, "stream ends after %llu of %llu samples", (unsigned long long)done, (unsigned long long)info.total_samples);
    if (check_md5) {
        uint8_t digest[16];
        md5_final(&md5, digest);
        if (memcmp(digest, info.md5, 16) != 0) BAIL(WB_AUDIO_ERR_CHECKSUM, "MD5 of the decoded PCM differs from STREAMINFO's signature");
    }
finish:
#undef BAIL
    free(chan);
    free(pcm_bytes);
    *decoded = rc ? 0 : done;
    return rc;
}
