/* Plain-C caller of libdeepfm_b200.so: proves include/deepfm_b200.h is consumable without Python, torch or C++.
 *
 *   gcc -std=c99 -I include integration/c_driver.c -o integration/_build/c_driver -L recommender_tensorflow_b200 -ldeepfm_b200
 *   c_driver --no-gpu     version string + argument validation only (what the CPU-only build check runs)
 *   c_driver              builds a small DeepFM (hashed int column, bucketized column, identity column, one numeric),
 *                         trains 40 steps from HOST buffers through dfm_train_step_host (H2D + step + D2H inside the call),
 *                         then evaluates; exits 0 when the loss fell and every status was DFM_OK.
 * The model mirrors what trainers/deep_fm.py:model_fn builds from get_feature_columns()-style descriptors
 * (trainers/ml_100k.py:18-39) - here spelled out with the C structs. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "deepfm_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc_ = (call);                                                                    \
        if (rc_ != DFM_OK) {                                                                 \
            fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, dfm_last_error(h));               \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(int argc, char** argv) {
    dfm_handle* h = NULL;
    printf("library: %s\n", dfm_version());
    {   /* trainers/deep_fm.py:31-34: no feature columns -> error, before any CUDA call */
        dfm_config bad;
        memset(&bad, 0, sizeof bad);
        if (dfm_create(&bad, &h) != DFM_ERR_INVALID_ARG || !strstr(dfm_last_error(NULL), "At least 1 feature column")) {
            fprintf(stderr, "argument validation failed\n");
            return 1;
        }
    }
    if (argc > 1 && !strcmp(argv[1], "--no-gpu")) { printf("C_DRIVER_OK (no gpu)\n"); return 0; }

    const float age_bounds[] = {15.f, 25.f, 35.f, 45.f, 55.f, 65.f};
    dfm_column cols[3];
    memset(cols, 0, sizeof cols);
    cols[0].name = "age_bucketized"; cols[0].kind = DFM_COL_BUCKETIZED; cols[0].dtype = DFM_INT32; cols[0].boundaries = age_bounds; cols[0].n_boundaries = 6;
    cols[1].name = "flag"; cols[1].kind = DFM_COL_IDENTITY; cols[1].dtype = DFM_INT32; cols[1].num_buckets = 2;
    cols[2].name = "user_id"; cols[2].kind = DFM_COL_HASH; cols[2].dtype = DFM_INT32; cols[2].num_buckets = 1000;
    const int32_t hidden[] = {16, 16};
    dfm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_cat = 3; cfg.cat = cols; cfg.n_num = 1; cfg.embedding_size = 4; cfg.n_hidden = 2; cfg.hidden_units = hidden;
    cfg.use_linear = cfg.use_mf = cfg.use_dnn = 1; cfg.loss_reduction = DFM_LOSS_MEAN;
    cfg.opt_deep.kind = cfg.opt_linear.kind = DFM_OPT_ADAM;
    cfg.opt_deep.lr = cfg.opt_linear.lr = 0.01f;
    cfg.opt_deep.beta1 = cfg.opt_linear.beta1 = 0.9f; cfg.opt_deep.beta2 = cfg.opt_linear.beta2 = 0.999f;
    cfg.opt_deep.eps = cfg.opt_linear.eps = 1e-8f;
    cfg.max_batch = 512; cfg.device = 0; cfg.rank = 0; cfg.world = 1;
    CHECK(dfm_create(&cfg, &h));
    CHECK(dfm_init_random(h, 7));

    enum { B = 512 };
    static int32_t age[B], flag[B], user[B];
    static float x[B], y[B], logits[B];
    const void* cat_data[3] = {age, flag, user};
    const int32_t* cat_off[3] = {NULL, NULL, NULL};
    const float* num_data[1] = {x};
    dfm_raw_batch batch;
    batch.batch_size = B; batch.cat_data = cat_data; batch.cat_offsets = cat_off; batch.num_data = num_data; batch.labels = y;
    unsigned s = 12345u;
    float first = 0.f, loss = 0.f;
    for (int step = 0; step < 40; ++step) {
        for (int b = 0; b < B; ++b) {
            s = s * 1664525u + 1013904223u; age[b] = 7 + (int)((s >> 8) % 66);
            s = s * 1664525u + 1013904223u; flag[b] = (int)((s >> 8) & 1);
            s = s * 1664525u + 1013904223u; user[b] = 1 + (int)((s >> 8) % 943);
            s = s * 1664525u + 1013904223u; x[b] = (float)((s >> 8) % 1000) / 1000.f;
            y[b] = (flag[b] ^ (age[b] > 35)) ? 1.f : 0.f;          /* needs the feature interaction */
        }
        CHECK(dfm_train_step_host(h, &batch, &loss, NULL));
        if (step == 0) first = loss;
    }
    CHECK(dfm_forward_host(h, &batch, logits));
    CHECK(dfm_sync(h));
    int ok = isfinite(loss) && loss < 0.8f * first && dfm_global_step(h) == 40;
    printf("loss %.4f -> %.4f after %lld steps, logits[0] = %.4f\n", first, loss, (long long)dfm_global_step(h), logits[0]);
    dfm_destroy(h);
    printf(ok ? "C_DRIVER_OK\n" : "C_DRIVER_FAILED\n");
    return ok ? 0 : 1;
}
