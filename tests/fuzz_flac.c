/* Mutation fuzzer for the FLAC decoder (tests/test_flac.py builds it with -fsanitize=address,undefined and runs it):
 * bit flips, random bytes, truncations and runs of 0x00 / 0xff over valid streams; the decoder may return any status but must
 * never read or write out of bounds or hit undefined behaviour.  usage: fuzz_flac <iterations> file.flac ... */
#include "../include/wb_audio.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t g_state = 88172645463325252ull;
static uint64_t rnd(void) {
    g_state ^= g_state << 13;
    g_state ^= g_state >> 7;
    g_state ^= g_state << 17;
    return g_state;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    long iterations = atol(argv[1]), ok = 0, format = 0, checksum = 0, other = 0;
    for (int a = 2; a < argc; ++a) {
        FILE* f = fopen(argv[a], "rb");
        if (!f) return 2;
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        uint8_t* orig = malloc((size_t)n);
        if (fread(orig, 1, (size_t)n, f) != (size_t)n) return 2;
        fclose(f);
        for (long it = 0; it < iterations; ++it) {
            long m = n;
            uint8_t* d = malloc((size_t)n); /* exact-size heap block: the sanitizer sees any read past the end */
            memcpy(d, orig, (size_t)n);
            switch (rnd() % 4) {
                case 0: d[rnd() % n] ^= (uint8_t)(1u << (rnd() % 8)); break;
                case 1:
                    for (int j = 0, k = 1 + (int)(rnd() % 8); j < k; ++j) d[rnd() % n] = (uint8_t)rnd();
                    break;
                case 2: {
                    m = (long)(rnd() % n);
                    uint8_t* t = malloc(m ? (size_t)m : 1);
                    memcpy(t, d, (size_t)m);
                    free(d);
                    d = t;
                    break;
                }
                default:
                    for (long p = (long)(rnd() % n), len = 1 + (long)(rnd() % 64), j = p; j < n && j < p + len; ++j) d[j] = (rnd() & 1) ? 0xff : 0;
            }
            uint64_t capacity = 1 + rnd() % 8000, got = 0;
            int32_t* out = malloc(capacity * 8 * sizeof(int32_t));
            int rc = wb_flac_decode_i32(d, (size_t)m, (rnd() % 8) ? out : NULL, capacity, &got, (int)(it & 1));
            if (rc == WB_AUDIO_OK) ++ok;
            else if (rc == WB_AUDIO_ERR_FORMAT) ++format;
            else if (rc == WB_AUDIO_ERR_CHECKSUM) ++checksum;
            else ++other;
            free(out);
            free(d);
        }
        free(orig);
    }
    printf("ok %ld format %ld checksum %ld other %ld\n", ok, format, checksum, other);
    return 0;
}
