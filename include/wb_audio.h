/* wb_audio.h — host-side audio decoding for the scripts around the Whisper path (plain C ABI, no CUDA).
 *
 * The reference reads its evaluation audio through `datasets` -> soundfile/libsndfile: `ds[i]["audio"]["array"]` in
 * examples/whisper/run.py:266-267 (the bundled librispeech_asr_dummy table stores FLAC bytes).  Neither decoder is in this
 * image, so the FLAC stream decoder is part of this repo: libwb_audio.so, built from whisper_trtllm_b200/csrc/audio/ with gcc
 * and loaded with ctypes by whisper_trtllm_b200/audio.py (same convention as libwhisper_b200.so: status codes, message from
 * wb_audio_last_error(), caller-owned buffers, no hidden allocation across the boundary).
 *
 * Decoding is exact integer work; every frame's CRC-8 / CRC-16 is checked, and the MD5 signature of the decoded PCM stored
 * in STREAMINFO is verified on request, so a wrong decode cannot pass silently.
 */
#ifndef WB_AUDIO_H
#define WB_AUDIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(WB_AUDIO_BUILD)
#define WB_AUDIO_API __attribute__((visibility("default")))
#else
#define WB_AUDIO_API
#endif

enum {
    WB_AUDIO_OK = 0,
    WB_AUDIO_ERR_ARGUMENT = -1,    /* null pointer / buffer too small */
    WB_AUDIO_ERR_FORMAT = -2,      /* not a FLAC stream, reserved field, truncated data */
    WB_AUDIO_ERR_CHECKSUM = -3,    /* frame CRC or stream MD5 mismatch */
    WB_AUDIO_ERR_UNSUPPORTED = -4  /* valid FLAC outside what this decoder handles */
};

typedef struct wb_flac_info {
    uint32_t sample_rate;
    uint32_t channels;         /* 1..8 */
    uint32_t bits_per_sample;  /* 4..32 */
    uint32_t min_blocksize, max_blocksize;
    uint64_t total_samples;    /* per channel; 0 = unknown */
    uint64_t audio_offset;     /* byte offset of the first frame */
    uint8_t md5[16];           /* MD5 of the unencoded PCM (all zero = not set) */
} wb_flac_info;

WB_AUDIO_API const char* wb_audio_last_error(void);
WB_AUDIO_API int wb_audio_version(void);

/* Parses the stream marker and the metadata blocks (an ID3v2 tag in front is skipped). */
WB_AUDIO_API int wb_flac_read_info(const uint8_t* data, size_t size, wb_flac_info* info);

/* Decodes the whole stream into interleaved int32 PCM (sample i of channel c at out[i * channels + c], native sample values,
 * not scaled).  capacity = samples per channel the buffer can hold; *decoded = samples per channel written.
 * out == NULL: nothing is stored, the stream is decoded and checked and *decoded is its length (for streams whose
 * STREAMINFO does not record it).
 * verify_md5 != 0: the MD5 of the decoded PCM must equal STREAMINFO's signature when one is set. */
WB_AUDIO_API int wb_flac_decode_i32(const uint8_t* data, size_t size, int32_t* out, uint64_t capacity, uint64_t* decoded,
                                    int verify_md5);

/* MD5 of a byte buffer (exposed for the tests of the signature check). */
WB_AUDIO_API int wb_md5(const uint8_t* data, size_t size, uint8_t digest[16]);

#ifdef __cplusplus
}
#endif
#endif
