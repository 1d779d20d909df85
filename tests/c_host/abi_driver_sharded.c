/* Two GPUs driven from ONE plain-C process through the expert-sharded entry points of include/deeprest_b200.h:
 * dr_create (rank r on device r) / dr_load_weights / dr_comm_init / dr_comm_attach (arena pointers: same process) /
 * dr_forward_sharded — the calls a cgo host makes for SURVEY §8e; no NCCL, no torch, no host-side collective.
 * usage: abi_driver_sharded <blob.bin> <x.bin> <out.bin> F M B T world
 * Every rank ends up with the stacked forecasts [B,T,M,Q] on its device; they must be bit-identical across ranks (the
 * partial sums are added in rank order) and are written out for the comparison with the oracle.  The few CUDA runtime calls
 * (device selection, staging x, reading the result) are declared by hand: the driver needs no CUDA headers. */
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
    if (argc != 9) { fprintf(stderr, "usage: %s blob x out F M B T world\n", argv[0]); return 2; }
    int F = atoi(argv[4]), M = atoi(argv[5]), B = atoi(argv[6]), T = atoi(argv[7]), world = atoi(argv[8]);
    if (world < 2 || world > 8) return 2;
    size_t per_expert = 256 + (size_t)(128 * F + F) + 2 * ((size_t)384 * F + 384 * 128 + 768) + 1539;
    size_t nblob = per_expert * M, nx = (size_t)B * T * F, no = (size_t)B * T * M * DR_Q;
    float* blob = slurp(argv[1], nblob);
    float* x = slurp(argv[2], nx);
    float* out = (float*)calloc(no, sizeof(float));
    dr_model* m[8];
    void* arena[8];
    for (int r = 0; r < world; ++r) {
        dr_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.F = F; cfg.M = M; cfg.H = DR_H; cfg.Q = DR_Q;
        cfg.quantiles[0] = 0.05f; cfg.quantiles[1] = 0.50f; cfg.quantiles[2] = 0.95f;
        cfg.dropout_p = 0.5f; cfg.engine = DR_ENGINE_AUTO; cfg.device = r; cfg.rank = r; cfg.world = world; cfg.dtype = DR_DTYPE_F32;
        if (dr_create(&cfg, &m[r]) != DR_OK) { fprintf(stderr, "dr_create[%d]: %s\n", r, dr_last_error(NULL)); return 1; }
        if (dr_load_weights(m[r], blob, nblob) != DR_OK || dr_comm_init(m[r], B, T, NULL, &arena[r]) != DR_OK) {
            fprintf(stderr, "rank %d: %s\n", r, dr_last_error(m[r]));
            return 1;
        }
    }
    for (int r = 0; r < world; ++r)
        if (dr_comm_attach(m[r], NULL, arena) != DR_OK) { fprintf(stderr, "attach %d: %s\n", r, dr_last_error(m[r])); return 1; }
    /* the device entry point is asynchronous, so one host thread can issue every rank's forward; the host entry point
     * blocks until its rank's columns are in `out`, which needs the peers' partial sums: issue the peers first (device
     * variant, x staged by a first host call is not needed — use the host variant on the LAST rank only after the others run) */
    for (int pass = 0; pass < 3; ++pass) {                 /* three passes: buffer parity and epoch flags are exercised */
        float* dev[8];
        /* dr_forward_sharded_dev is asynchronous: one host thread issues every rank, then synchronises (a blocking host
         * entry point per rank would need one thread per rank, because a rank only finishes once its peers have sent
         * their partial sums) */
        extern int cudaSetDevice(int);
        extern int cudaMalloc(void**, size_t);
        extern int cudaMemcpy(void*, const void*, size_t, int);
        extern int cudaDeviceSynchronize(void);
        static float* xd[8];
        for (int r = 0; r < world; ++r) {
            cudaSetDevice(r);
            if (!xd[r]) cudaMalloc((void**)&xd[r], nx * sizeof(float));
            cudaMemcpy(xd[r], x, nx * sizeof(float), 1 /* cudaMemcpyHostToDevice */);
        }
        for (int r = 0; r < world; ++r)
            if (dr_forward_sharded_dev(m[r], xd[r], B, T, &dev[r]) != DR_OK) { fprintf(stderr, "forward %d: %s\n", r, dr_last_error(m[r])); return 1; }
        for (int r = 0; r < world; ++r) { cudaSetDevice(r); cudaDeviceSynchronize(); }
        /* every rank's stacked tensor must be the same bits: compare all of them with rank 0's */
        float* ref = (float*)malloc(no * sizeof(float));
        float* cur = (float*)malloc(no * sizeof(float));
        cudaSetDevice(0);
        cudaMemcpy(ref, dev[0], no * sizeof(float), 2 /* cudaMemcpyDeviceToHost */);
        for (int r = 1; r < world; ++r) {
            cudaSetDevice(r);
            cudaMemcpy(cur, dev[r], no * sizeof(float), 2);
            if (memcmp(ref, cur, no * sizeof(float)) != 0) { fprintf(stderr, "pass %d: rank %d holds different forecasts than rank 0\n", pass, r); return 1; }
        }
        memcpy(out, ref, no * sizeof(float));
        free(ref); free(cur);
    }
    FILE* f = fopen(argv[3], "wb");
    fwrite(out, sizeof(float), no, f);
    fclose(f);
    printf("engine=%s world=%d launches=%lld\n", dr_last_engine(m[0]), world, (long long)dr_launch_count(m[0]));
    for (int r = 0; r < world; ++r) dr_destroy(m[r]);
    free(blob); free(x); free(out);
    return 0;
}
