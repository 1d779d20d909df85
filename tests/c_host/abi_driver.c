/* A non-Python host of libdeeprest_b200.so: plain C, only include/deeprest_b200.h.
 * usage: abi_driver <blob.bin> <x.bin> <out.bin> F M B T
 * Reads the weight blob and the windows as raw little-endian fp32, runs dr_create / dr_load_weights / dr_forward /
 * dr_quantile_loss against labels == 0, writes the forecasts.  tests/test_gpu_c_host.py builds it with gcc and
 * compares the output with the oracle — the same calls a cgo shim would make. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "deeprest_b200.h"

static float* slurp(const char* path, size_t n) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    float* p = (float*)malloc(n * sizeof(float));
    if (fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "%s: short read\n", path); exit(2); }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage: %s blob x out F M B T\n", argv[0]); return 2; }
    int F = atoi(argv[4]), M = atoi(argv[5]), B = atoi(argv[6]), T = atoi(argv[7]);
    dr_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.F = F; cfg.M = M; cfg.H = DR_H; cfg.Q = DR_Q;
    cfg.quantiles[0] = 0.05f; cfg.quantiles[1] = 0.50f; cfg.quantiles[2] = 0.95f;
    cfg.dropout_p = 0.5f; cfg.engine = DR_ENGINE_AUTO; cfg.device = 0; cfg.rank = 0; cfg.world = 1;
    dr_model* m = NULL;
    if (dr_create(&cfg, &m) != DR_OK) { fprintf(stderr, "dr_create: %s\n", dr_last_error(NULL)); return 1; }
    size_t per_expert = 256 + (size_t)(128 * F + F) + 2 * ((size_t)384 * F + 384 * 128 + 768) + 1539;   /* SURVEY §8 P_e */
    size_t nblob = per_expert * M, nx = (size_t)B * T * F, no = (size_t)B * T * M * DR_Q;
    float* blob = slurp(argv[1], nblob);
    float* x = slurp(argv[2], nx);
    float* out = (float*)malloc(no * sizeof(float));
    float* y = (float*)calloc((size_t)B * T * M, sizeof(float));
    float loss = -1.0f;
    if (dr_load_weights(m, blob, nblob) != DR_OK || dr_forward(m, x, B, T, out) != DR_OK ||
        dr_quantile_loss(m, out, y, B, T, &loss) != DR_OK) {
        fprintf(stderr, "deeprest: %s\n", dr_last_error(m));
        return 1;
    }
    /* error path: a forward with a bad shape must fail cleanly and leave a message */
    if (dr_forward(m, x, 0, T, out) != DR_EINVAL || strlen(dr_last_error(m)) == 0) { fprintf(stderr, "bad-shape check failed\n"); return 1; }
    FILE* f = fopen(argv[3], "wb");
    fwrite(out, sizeof(float), no, f);
    fclose(f);
    printf("engine=%s launches=%lld loss=%.7f\n", dr_last_engine(m), (long long)dr_launch_count(m), loss);
    dr_destroy(m);
    free(blob); free(x); free(out); free(y);
    return 0;
}
