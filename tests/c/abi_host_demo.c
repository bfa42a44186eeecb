/* Plain-C host of liblighthand_b200.so: binds include/lighthand_b200.h exactly as a non-Python maintainer
 * would, runs one step through the host-buffer context and prints the results as JSON for the test to check.
 *   gcc -std=c11 -O2 -I include tests/c/abi_host_demo.c -o demo -L lighthand_b200 -llighthand_b200 -lm
 * Input: argv[1] = raw file with  images u8[N,S,S,3] | affine f64[N,6] | flip u8[N] | joints f32[N,21,2] | pred f32[N,21,H,H]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lighthand_b200.h"

#define NB 8   /* batch */
#define SRC 64 /* source = network input side */
#define HM 16  /* heat-map side */
#define NJ 21  /* joints */

static void die(const char* what, int rc) {
    fprintf(stderr, "%s failed: [%d] %s\n", what, rc, lh_strerror(rc));
    exit(1);
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (lh_abi_version() != LH_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 3; }
    const size_t n_img = (size_t)NB * SRC * SRC * 3, n_pred = (size_t)NB * NJ * HM * HM;
    unsigned char* images = malloc(n_img);
    double* affine = malloc(sizeof(double) * NB * 6);
    unsigned char* flip = malloc(NB);
    float* joints = malloc(sizeof(float) * NB * NJ * 2);
    float* pred = malloc(sizeof(float) * n_pred);
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 4;
    if (fread(images, 1, n_img, f) != n_img || fread(affine, sizeof(double), NB * 6, f) != NB * 6 ||
        fread(flip, 1, NB, f) != NB || fread(joints, sizeof(float), NB * NJ * 2, f) != NB * NJ * 2 ||
        fread(pred, sizeof(float), n_pred, f) != n_pred) return 5;
    fclose(f);

    /* the 13x13 table exactly as the reference computes it: float32 exp of a float32 argument */
    float table[13 * 13];
    for (int y = 0; y < 13; ++y)
        for (int x = 0; x < 13; ++x) {
            const float a = (float)((x - 6) * (x - 6) + (y - 6) * (y - 6));
            table[y * 13 + x] = expf(-a / 8.0f);
        }

    lh_pipeline_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = 0; cfg.batch = NB; cfg.src_h = SRC; cfg.src_w = SRC; cfg.out_h = SRC; cfg.out_w = SRC;
    cfg.J = NJ; cfg.H = HM; cfg.W = HM; cfg.interp = LH_INTERP_CV2_U8; cfg.train = 1; cfg.materialize_target = 0;
    cfg.exact_norm = 1; cfg.sigma = 2.0f; cfg.pck_T = 0.2f;
    const float ms[6] = {0.485f, 0.456f, 0.406f, 0.229f, 0.224f, 0.225f};
    memcpy(cfg.mean_std, ms, sizeof ms);

    lh_pipeline* p = NULL;
    int rc = lh_pipeline_create(&cfg, table, 13, &p);
    if (rc) die("lh_pipeline_create", rc);

    float loss = 0.f;
    float* coords = malloc(sizeof(float) * NB * NJ * 2);
    float* maxvals = malloc(sizeof(float) * NB * NJ);
    unsigned long long counts[4] = {0, 0, 0, 0};
    double sums[4] = {0, 0, 0, 0};
    lh_pipeline_io io;
    memset(&io, 0, sizeof io);
    io.images = images; io.fwd_affine = affine; io.flip = flip; io.joints = joints; io.pred = pred;
    io.loss = &loss; io.coords = coords; io.maxvals = maxvals; io.counts = counts; io.sums = sums;
    for (int rep = 0; rep < 2; ++rep) {           /* two steps (a third one below): the counters accumulate */
        int slot = -1;
        rc = lh_pipeline_submit(p, &io, &slot);
        if (rc) die("lh_pipeline_submit", rc);
        rc = lh_pipeline_wait(p, slot);
        if (rc) die("lh_pipeline_wait", rc);
    }
    /* the split form of a step: data stage, (backbone on the loss stream), loss stage */
    {
        int slot = -1;
        rc = lh_pipeline_submit_inputs(p, &io, &slot);
        if (rc) die("lh_pipeline_submit_inputs", rc);
        if (!lh_pipeline_loss_stream(p) || !lh_pipeline_input_dev_slot(p, slot)) die("loss stream / input slot", -2);
        rc = lh_pipeline_submit_loss(p, slot);
        if (rc) die("lh_pipeline_submit_loss", rc);
        rc = lh_pipeline_wait(p, slot);
        if (rc) die("lh_pipeline_wait (split)", rc);
    }
    /* K6 at world size 1: own communicator, all-reduce of the accumulators = identity */
    int allreduce_ok = 0;
    {
        unsigned char id[LH_COMM_ID_BYTES];
        void* comm = NULL;
        unsigned long long c2[4];
        double s2[4];
        rc = lh_comm_unique_id(id);
        if (rc) die("lh_comm_unique_id", rc);
        rc = lh_comm_init(id, 1, 0, 0, &comm);
        if (rc) die("lh_comm_init", rc);
        rc = lh_pipeline_allreduce_metrics(p, comm);
        if (rc) die("lh_pipeline_allreduce_metrics", rc);
        rc = lh_pipeline_read_metrics(p, c2, s2);
        if (rc) die("lh_pipeline_read_metrics", rc);
        allreduce_ok = (memcmp(c2, counts, sizeof c2) == 0) && (memcmp(s2, sums, sizeof s2) == 0);
        rc = lh_comm_destroy(comm);
        if (rc) die("lh_comm_destroy", rc);
    }
    /* bad arguments are reported, not crashed on */
    const int e1 = lh_pipeline_submit(p, NULL, NULL);
    const int e2 = lh_decode(NULL, 1, 21, 64, 64, 4.0f, 0, NULL, NULL, NULL, NULL, NULL);

    double csum = 0.0;
    for (int i = 0; i < NB * NJ * 2; ++i) csum += coords[i];
    printf("{\"loss\": %.9g, \"coords_sum\": %.9g, \"coord0\": [%.9g, %.9g], \"maxval0\": %.9g, "
           "\"counts\": [%llu, %llu, %llu, %llu], \"epe_sum\": %.12g, \"launches\": %d, \"e1\": %d, \"e2\": %d, "
           "\"allreduce_ok\": %d}\n",
           loss, csum, coords[0], coords[1], maxvals[0], counts[0], counts[1], counts[2], counts[3], sums[0],
           lh_pipeline_launches_per_step(p), e1, e2, allreduce_ok);
    lh_pipeline_destroy(p);
    free(images); free(affine); free(flip); free(joints); free(pred); free(coords); free(maxvals);
    return 0;
}
