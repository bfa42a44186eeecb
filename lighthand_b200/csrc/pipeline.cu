// Host-buffer pipeline context: what a host without torch binds, and the end-to-end path
// (host u8 images + matrices + joints in, loss / coordinates / metric counts out).
//
// Two slots, three streams, ordering by CUDA events only (the host blocks once per step on the slot's "done"
// event):
//   copy stream   H2D of step i+1 (DMA engine)                    } overlap the
//   data stream   K1 warp+normalise of step i+1 -> input[slot]    } loss stage of step i
//   loss stream   [the caller's backbone] -> fused K2+K3+K4+K5 of step i -> D2H of the results
// which is the shape of the reference's loop (src/utils/method.py:160-183): the data stage of the next batch runs
// ahead of the model / loss stage of the current one.  K1 and the fused pass touch disjoint buffers, so they may
// run concurrently; each slot owns its network-input buffer.  Kernels per step: 2 launches (the normalisation
// table is built once at create time).
#include <cstdlib>
#include <cstring>
#include <new>

#include <thread>

#include "lh_common.cuh"

namespace {

constexpr int kSlots = 2;

struct Slot {
    uint8_t* d_images = nullptr;
    double* d_affine = nullptr;
    uint8_t* d_flip = nullptr;
    float* d_joints = nullptr;
    float* d_coords = nullptr;
    float* d_maxvals = nullptr;
    float* d_loss = nullptr;
    float* d_input = nullptr;   // this slot's network input [N,3,out_h,out_w] (K1 of the next step must not overwrite it)
    // pinned result block: loss[1] | coords[N*J*2] | maxvals[N*J] | counts u64[4] | sums f64[4]
    unsigned char* h_out = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    bool busy = false;        // submitted (inputs at least), not yet waited
    bool loss_pending = false; // lh_pipeline_submit_inputs done, lh_pipeline_submit_loss not yet
    lh_pipeline_io io{};
};

}  // namespace

struct lh_pipeline {
    lh_pipeline_cfg cfg{};
    int table_side = 0;
    cudaStream_t s_copy = nullptr, s_data = nullptr, s_comp = nullptr;   // s_comp = the loss stream
    Slot slot[kSlots];
    int last_slot = 0;
    float* d_pred = nullptr;
    float* d_target = nullptr;
    float* d_grad = nullptr;
    float* d_table = nullptr;
    float* d_lut = nullptr;
    void* d_ws = nullptr;
    unsigned long long* d_counts = nullptr;
    double* d_sums = nullptr;
    size_t out_bytes = 0, off_coords = 0, off_maxvals = 0, off_counts = 0, off_sums = 0;
    unsigned long long step = 0;
    lh_jpeg* jpeg = nullptr;    // created by the first lh_pipeline_submit_inputs_jpeg
    int jpeg_backend = -1;
};

#define PCHK(x)                                      \
    do {                                             \
        cudaError_t e__ = (x);                       \
        if (e__ != cudaSuccess) { rc = (int)e__; goto fail; } \
    } while (0)

extern "C" void lh_pipeline_destroy(lh_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->cfg.device);
    if (p->s_comp) cudaStreamSynchronize(p->s_comp);
    if (p->s_data) cudaStreamSynchronize(p->s_data);
    if (p->s_copy) cudaStreamSynchronize(p->s_copy);
    for (auto& s : p->slot) {
        cudaFree(s.d_images); cudaFree(s.d_affine); cudaFree(s.d_flip); cudaFree(s.d_joints);
        cudaFree(s.d_coords); cudaFree(s.d_maxvals); cudaFree(s.d_loss); cudaFree(s.d_input);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
    }
    cudaFree(p->d_pred); cudaFree(p->d_target); cudaFree(p->d_grad);
    cudaFree(p->d_table); cudaFree(p->d_lut); cudaFree(p->d_ws); cudaFree(p->d_counts); cudaFree(p->d_sums);
    if (p->jpeg) lh_jpeg_destroy(p->jpeg);
    if (p->s_copy) cudaStreamDestroy(p->s_copy);
    if (p->s_data) cudaStreamDestroy(p->s_data);
    if (p->s_comp) cudaStreamDestroy(p->s_comp);
    delete p;
}

extern "C" int lh_pipeline_create(const lh_pipeline_cfg* cfg, const float* table_host, int table_side,
                                  lh_pipeline** out) {
    if (!cfg || !table_host || !out) return LH_ENULL;
    if (cfg->batch <= 0 || cfg->src_h <= 0 || cfg->src_w <= 0 || cfg->out_h <= 0 || cfg->out_w <= 0 ||
        cfg->J <= 0 || cfg->H <= 0 || cfg->W <= 0 || table_side < 1 || (table_side & 1) == 0 || table_side > 31)
        return LH_EINVAL;
    if (cfg->interp < LH_INTERP_CV2_U8 || cfg->interp > LH_INTERP_TORCH) return LH_EINVAL;
    int rc = LH_OK;
    lh_pipeline* p = new (std::nothrow) lh_pipeline();
    if (!p) return LH_ENOMEM;
    p->cfg = *cfg;
    p->table_side = table_side;
    const size_t N = cfg->batch, J = cfg->J, HW = (size_t)cfg->H * cfg->W;
    const size_t img_bytes = (size_t)cfg->src_h * cfg->src_w * 3;
    PCHK(cudaSetDevice(cfg->device));
    PCHK(cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking));
    PCHK(cudaStreamCreateWithFlags(&p->s_data, cudaStreamNonBlocking));
    PCHK(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
    p->off_coords = 16;
    p->off_maxvals = p->off_coords + N * J * 2 * sizeof(float);
    p->off_counts = (p->off_maxvals + N * J * sizeof(float) + 15) & ~(size_t)15;
    p->off_sums = p->off_counts + 4 * sizeof(unsigned long long);
    p->out_bytes = p->off_sums + 4 * sizeof(double);
    for (auto& s : p->slot) {
        PCHK(cudaMalloc(&s.d_images, N * img_bytes));
        PCHK(cudaMalloc(&s.d_affine, N * 6 * sizeof(double)));
        PCHK(cudaMalloc(&s.d_flip, N));
        PCHK(cudaMemset(s.d_flip, 0, N));
        PCHK(cudaMalloc(&s.d_joints, N * J * 2 * sizeof(float)));
        PCHK(cudaMalloc(&s.d_coords, N * J * 2 * sizeof(float)));
        PCHK(cudaMalloc(&s.d_maxvals, N * J * sizeof(float)));
        PCHK(cudaMalloc(&s.d_loss, 16));
        PCHK(cudaMalloc(&s.d_input, N * 3 * (size_t)cfg->out_h * cfg->out_w * sizeof(float)));
        PCHK(cudaMallocHost(&s.h_out, p->out_bytes));
        PCHK(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        PCHK(cudaEventCreateWithFlags(&s.ev_k1, cudaEventDisableTiming));
        PCHK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    PCHK(cudaMalloc(&p->d_pred, N * J * HW * sizeof(float)));
    PCHK(cudaMemset(p->d_pred, 0, N * J * HW * sizeof(float)));
    if (cfg->materialize_target) PCHK(cudaMalloc(&p->d_target, N * J * HW * sizeof(float)));
    if (cfg->train) PCHK(cudaMalloc(&p->d_grad, N * J * HW * sizeof(float)));
    PCHK(cudaMalloc(&p->d_table, (size_t)table_side * table_side * sizeof(float)));
    PCHK(cudaMemcpy(p->d_table, table_host, (size_t)table_side * table_side * sizeof(float), cudaMemcpyHostToDevice));
    PCHK(cudaMalloc(&p->d_lut, 768 * sizeof(float)));
    {
        const size_t wsb = lh_heatmap_ws_bytes((int64_t)N, (int)J);
        PCHK(cudaMalloc(&p->d_ws, wsb));
        PCHK(cudaMemset(p->d_ws, 0, wsb));
    }
    PCHK(cudaMalloc(&p->d_counts, 4 * sizeof(unsigned long long)));
    PCHK(cudaMalloc(&p->d_sums, 4 * sizeof(double)));
    PCHK(cudaMemset(p->d_counts, 0, 4 * sizeof(unsigned long long)));
    PCHK(cudaMemset(p->d_sums, 0, 4 * sizeof(double)));
    rc = lh_norm_lut_build(nullptr, 1, cfg->mean_std, p->d_lut, p->s_comp);   // used when cfg.exact_norm
    if (rc) goto fail;
    PCHK(cudaStreamSynchronize(p->s_comp));
    *out = p;
    return LH_OK;
fail:
    lh_pipeline_destroy(p);
    return rc;
}

extern "C" float* lh_pipeline_input_dev(lh_pipeline* p) { return p ? p->slot[p->last_slot].d_input : nullptr; }
extern "C" float* lh_pipeline_input_dev_slot(lh_pipeline* p, int slot) {
    return (p && slot >= 0 && slot < kSlots) ? p->slot[slot].d_input : nullptr;
}
extern "C" void* lh_pipeline_loss_stream(lh_pipeline* p) { return p ? (void*)p->s_comp : nullptr; }
extern "C" float* lh_pipeline_pred_dev(lh_pipeline* p) { return p ? p->d_pred : nullptr; }
extern "C" float* lh_pipeline_grad_dev(lh_pipeline* p) { return p ? p->d_grad : nullptr; }
extern "C" float* lh_pipeline_target_dev(lh_pipeline* p) { return p ? p->d_target : nullptr; }
extern "C" int lh_pipeline_launches_per_step(const lh_pipeline*) { return 2; }

extern "C" int lh_pipeline_reset_metrics(lh_pipeline* p) {
    if (!p) return LH_ENULL;
    cudaSetDevice(p->cfg.device);
    cudaError_t e = cudaMemsetAsync(p->d_counts, 0, 4 * sizeof(unsigned long long), p->s_comp);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_sums, 0, 4 * sizeof(double), p->s_comp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_comp);
    return (int)e;
}

// data stage of one step into slot (step % 2): H2D on the copy stream, K1 on the data stream; the loss stream is
// made to wait for K1, so whatever the caller enqueues on lh_pipeline_loss_stream() next (the backbone) sees the input
static int submit_inputs_impl(lh_pipeline* p, const lh_pipeline_io* io, const uint8_t* const* jpeg_data,
                              const size_t* jpeg_sizes, int* slot_out) {
    if (!p || !io || !(io->images || jpeg_data) || !io->joints || !slot_out) return LH_ENULL;
    const lh_pipeline_cfg& c = p->cfg;
    if (c.interp != LH_INTERP_TORCH && !io->fwd_affine) return LH_ENULL;
    const int si = (int)(p->step % kSlots);
    Slot& s = p->slot[si];
    if (s.busy) return LH_EINVAL;  // caller must lh_pipeline_wait() the slot first
    const size_t N = c.batch, J = c.J;
    const size_t img_bytes = (size_t)c.src_h * c.src_w * 3;
    int rc = LH_OK;
    PCHK(cudaSetDevice(c.device));
    // H2D on the copy stream (the slot's previous consumer finished: its ev_done was waited by the host)
    if (jpeg_data) {
        // compressed input: PCIe carries the JPEG bytes, nvJPEG decodes straight into the slot's uint8 HWC source buffer
        rc = lh_jpeg_decode_batch(p->jpeg, jpeg_data, jpeg_sizes, (int)N, c.src_w, c.src_h, s.d_images, (int64_t)img_bytes,
                                  c.src_w * 3, p->s_copy);
        if (rc) goto fail;
    } else {
        PCHK(cudaMemcpyAsync(s.d_images, io->images, N * img_bytes, cudaMemcpyHostToDevice, p->s_copy));
    }
    if (io->fwd_affine) PCHK(cudaMemcpyAsync(s.d_affine, io->fwd_affine, N * 6 * sizeof(double), cudaMemcpyHostToDevice, p->s_copy));
    if (io->flip) PCHK(cudaMemcpyAsync(s.d_flip, io->flip, N, cudaMemcpyHostToDevice, p->s_copy));
    PCHK(cudaMemcpyAsync(s.d_joints, io->joints, N * J * 2 * sizeof(float), cudaMemcpyHostToDevice, p->s_copy));
    PCHK(cudaEventRecord(s.ev_h2d, p->s_copy));
    PCHK(cudaStreamWaitEvent(p->s_data, s.ev_h2d, 0));
    rc = lh_warp_normalize(s.d_images, (int64_t)img_bytes, c.src_h, c.src_w, c.src_w * 3, s.d_affine,
                           io->flip ? s.d_flip : nullptr, nullptr, c.mean_std,
                           (c.exact_norm || c.interp != LH_INTERP_CV2_U8) ? p->d_lut : nullptr, c.interp, LH_CHAN_RGB,
                           LH_LAYOUT_CHW, s.d_input, (int)N, c.out_h, c.out_w, p->s_data);
    if (rc) goto fail;
    PCHK(cudaEventRecord(s.ev_k1, p->s_data));
    PCHK(cudaStreamWaitEvent(p->s_comp, s.ev_k1, 0));
    s.io = *io;
    s.busy = true;
    s.loss_pending = true;
    p->last_slot = si;
    p->step++;
    *slot_out = si;
    return LH_OK;
fail:
    return rc;
}

extern "C" int lh_pipeline_submit_inputs(lh_pipeline* p, const lh_pipeline_io* io, int* slot_out) {
    if (!io || !io->images) return LH_ENULL;
    return submit_inputs_impl(p, io, nullptr, nullptr, slot_out);
}

// the same data stage fed with N JPEG bit-streams (io->images is ignored): decode on the device, then K1
extern "C" int lh_pipeline_submit_inputs_jpeg(lh_pipeline* p, const uint8_t* const* data, const size_t* sizes,
                                              int jpeg_backend, const lh_pipeline_io* io, int* slot_out) {
    if (!p || !data || !sizes) return LH_ENULL;
    if (p->jpeg && p->jpeg_backend != jpeg_backend) {       // another backend asked for: start over
        cudaSetDevice(p->cfg.device);
        cudaStreamSynchronize(p->s_copy);
        lh_jpeg_destroy(p->jpeg);
        p->jpeg = nullptr;
    }
    if (!p->jpeg) {
        // HYBRID decodes the entropy-coded data on the host: give it the host's cores (LH_JPEG_THREADS overrides, at most 32)
        int threads = 1;
        if (jpeg_backend == LH_JPEG_BACKEND_HYBRID) {
            const char* e = getenv("LH_JPEG_THREADS");
            threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
            threads = threads < 1 ? 1 : (threads > 32 ? 32 : threads);
        }
        const int rc = lh_jpeg_create(p->cfg.device, jpeg_backend, threads, &p->jpeg);
        if (rc) return rc;
        p->jpeg_backend = jpeg_backend;
    }
    return submit_inputs_impl(p, io, data, sizes, slot_out);
}

// loss stage of the step in `si` on the loss stream: [pred H2D] -> fused K2+K3+K4+K5 -> D2H of the results
extern "C" int lh_pipeline_submit_loss(lh_pipeline* p, int si) {
    if (!p) return LH_ENULL;
    if (si < 0 || si >= kSlots) return LH_EINVAL;
    Slot& s = p->slot[si];
    if (!s.busy || !s.loss_pending) return LH_EINVAL;
    const lh_pipeline_cfg& c = p->cfg;
    const lh_pipeline_io* io = &s.io;
    const size_t N = c.batch, J = c.J, HW = (size_t)c.H * c.W;
    int rc = LH_OK;
    PCHK(cudaSetDevice(c.device));
    if (io->pred)   // d_pred is shared by both slots: the loss stream is in order, so this copy follows the previous step's kernels
        PCHK(cudaMemcpyAsync(p->d_pred, io->pred, N * J * HW * sizeof(float), cudaMemcpyHostToDevice, p->s_comp));
    {
        lh_heatmap_args a{};
        a.pred = p->d_pred; a.joints = s.d_joints; a.gauss_table = p->d_table;
        a.target_out = p->d_target; a.grad = p->d_grad;
        a.loss_f32 = s.d_loss; a.loss_meter = p->d_sums + 1; a.coords = s.d_coords; a.maxvals = s.d_maxvals; a.ws = p->d_ws;
        a.N = (int64_t)N; a.J = c.J; a.H = c.H; a.W = c.W; a.joint_stride = 2; a.table_side = p->table_side;
        a.style = LH_TARGET_HRNET; a.use_target_weight = 0; a.subpixel = c.subpixel;
        a.stride_x = (float)c.out_w / (float)c.W; a.stride_y = (float)c.out_h / (float)c.H;
        a.out_scale = a.stride_x;
        a.metric_counts = p->d_counts; a.metric_sums = p->d_sums;      // K5 folded into the same launch
        a.pck_thr = c.pck_T; a.pck_mode = LH_PCK_PROPORTION;
        rc = lh_heatmap_fused(&a, p->s_comp);
        if (rc) goto fail;
    }
    PCHK(cudaMemcpyAsync(s.h_out, s.d_loss, sizeof(float), cudaMemcpyDeviceToHost, p->s_comp));
    if (io->coords) PCHK(cudaMemcpyAsync(s.h_out + p->off_coords, s.d_coords, N * J * 2 * sizeof(float), cudaMemcpyDeviceToHost, p->s_comp));
    if (io->maxvals) PCHK(cudaMemcpyAsync(s.h_out + p->off_maxvals, s.d_maxvals, N * J * sizeof(float), cudaMemcpyDeviceToHost, p->s_comp));
    PCHK(cudaMemcpyAsync(s.h_out + p->off_counts, p->d_counts, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->s_comp));
    PCHK(cudaMemcpyAsync(s.h_out + p->off_sums, p->d_sums, 4 * sizeof(double), cudaMemcpyDeviceToHost, p->s_comp));
    PCHK(cudaEventRecord(s.ev_done, p->s_comp));
    s.loss_pending = false;
    return LH_OK;
fail:
    return rc;
}

// one whole step (data stage + loss stage) into slot (step % 2)
extern "C" int lh_pipeline_submit(lh_pipeline* p, const lh_pipeline_io* io, int* slot_out) {
    const int rc = lh_pipeline_submit_inputs(p, io, slot_out);
    if (rc) return rc;
    return lh_pipeline_submit_loss(p, *slot_out);
}

// SUM the running metric accumulators over the ranks of `comm` (lh_comm_init), stream-ordered after the last
// submitted loss stage; every rank calls it the same number of times
extern "C" int lh_pipeline_allreduce_metrics(lh_pipeline* p, void* comm) {
    if (!p || !comm) return LH_ENULL;
    cudaError_t e = cudaSetDevice(p->cfg.device);
    if (e != cudaSuccess) return (int)e;
    return lh_metrics_allreduce(comm, p->d_counts, 4, p->d_sums, 4, p->s_comp);
}

// synchronous read of the running accumulators (after everything submitted so far)
extern "C" int lh_pipeline_read_metrics(lh_pipeline* p, unsigned long long* counts4, double* sums4) {
    if (!p || !counts4 || !sums4) return LH_ENULL;
    cudaError_t e = cudaSetDevice(p->cfg.device);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts4, p->d_counts, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->s_comp);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sums4, p->d_sums, 4 * sizeof(double), cudaMemcpyDeviceToHost, p->s_comp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_comp);
    return (int)e;
}

// block until slot `si` finished and hand its results to the host pointers given at submit time
extern "C" int lh_pipeline_wait(lh_pipeline* p, int si) {
    if (!p) return LH_ENULL;
    if (si < 0 || si >= kSlots) return LH_EINVAL;
    Slot& s = p->slot[si];
    if (!s.busy || s.loss_pending) return LH_EINVAL;   // nothing submitted, or the loss stage is still missing
    cudaSetDevice(p->cfg.device);
    cudaError_t e = cudaEventSynchronize(s.ev_done);
    s.busy = false;
    if (e != cudaSuccess) return (int)e;
    const size_t N = p->cfg.batch, J = p->cfg.J;
    if (s.io.loss) std::memcpy(s.io.loss, s.h_out, sizeof(float));
    if (s.io.coords) std::memcpy(s.io.coords, s.h_out + p->off_coords, N * J * 2 * sizeof(float));
    if (s.io.maxvals) std::memcpy(s.io.maxvals, s.h_out + p->off_maxvals, N * J * sizeof(float));
    if (s.io.counts) std::memcpy(s.io.counts, s.h_out + p->off_counts, 4 * sizeof(unsigned long long));
    if (s.io.sums) std::memcpy(s.io.sums, s.h_out + p->off_sums, 4 * sizeof(double));
    return LH_OK;
}

extern "C" int lh_pipeline_step_host(lh_pipeline* p, const lh_pipeline_io* io) {
    int si = -1;
    const int rc = lh_pipeline_submit(p, io, &si);
    if (rc) return rc;
    return lh_pipeline_wait(p, si);
}

// ------------------------------------------------------------------------------------------------
// Stage ring: the stream / event plumbing of the pipelined DEVICE-resident form (HeatmapPipeline.data_stage / loss_stage,
// PreparedStep.launch_data / launch_loss) in C, so that one step costs two library calls and no per-step Python stream or
// event work -- at ~90 us per step eight ranks sharing a host could not keep their GPUs fed through torch's wrappers.
// Two slots (network-input buffers owned by the caller); a data stream owned by the ring.
// ------------------------------------------------------------------------------------------------
struct lh_stage_ring {
    int device = 0;
    cudaStream_t s_data = nullptr;
    cudaEvent_t ev_ready[kSlots] = {nullptr, nullptr};   // K1 of the slot finished (recorded on the data stream)
    cudaEvent_t ev_free[kSlots] = {nullptr, nullptr};    // the slot's consumer finished (recorded on the caller's stream)
    cudaEvent_t ev_in = nullptr;                         // "inputs produced so far on the caller's stream"
    bool free_valid[kSlots] = {false, false};
};

extern "C" void lh_stage_ring_destroy(lh_stage_ring* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->s_data) cudaStreamSynchronize(r->s_data);
    for (int i = 0; i < kSlots; ++i) {
        if (r->ev_ready[i]) cudaEventDestroy(r->ev_ready[i]);
        if (r->ev_free[i]) cudaEventDestroy(r->ev_free[i]);
    }
    if (r->ev_in) cudaEventDestroy(r->ev_in);
    if (r->s_data) cudaStreamDestroy(r->s_data);
    delete r;
}

extern "C" int lh_stage_ring_create(int device, lh_stage_ring** out) {
    if (!out) return LH_ENULL;
    int rc = LH_OK;
    lh_stage_ring* r = new (std::nothrow) lh_stage_ring();
    if (!r) return LH_ENOMEM;
    r->device = device;
    PCHK(cudaSetDevice(device));
    PCHK(cudaStreamCreateWithFlags(&r->s_data, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
        PCHK(cudaEventCreateWithFlags(&r->ev_ready[i], cudaEventDisableTiming));
        PCHK(cudaEventCreateWithFlags(&r->ev_free[i], cudaEventDisableTiming));
    }
    PCHK(cudaEventCreateWithFlags(&r->ev_in, cudaEventDisableTiming));
    *out = r;
    return LH_OK;
fail:
    lh_stage_ring_destroy(r);
    return rc;
}

extern "C" void* lh_stage_ring_data_stream(lh_stage_ring* r) { return r ? (void*)r->s_data : nullptr; }

// data stream: wait for the inputs (an event, or everything enqueued so far on `after_stream`, or nothing) and for the
// previous consumer of the slot's input buffer
extern "C" int lh_stage_ring_begin_data(lh_stage_ring* r, int slot, void* wait_event, void* after_stream, int wait_mode) {
    if (!r) return LH_ENULL;
    if (slot < 0 || slot >= kSlots || wait_mode < 0 || wait_mode > 2) return LH_EINVAL;
    cudaError_t e = cudaSuccess;
    if (wait_mode == 1) {                     // an event the caller recorded
        if (!wait_event) return LH_ENULL;
        e = cudaStreamWaitEvent(r->s_data, (cudaEvent_t)wait_event, 0);
    } else if (wait_mode == 2) {              // everything enqueued so far on the caller's stream
        e = cudaEventRecord(r->ev_in, (cudaStream_t)after_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(r->s_data, r->ev_in, 0);
    }
    if (e == cudaSuccess && r->free_valid[slot]) e = cudaStreamWaitEvent(r->s_data, r->ev_free[slot], 0);
    return (int)e;
}
extern "C" int lh_stage_ring_end_data(lh_stage_ring* r, int slot) {
    if (!r) return LH_ENULL;
    if (slot < 0 || slot >= kSlots) return LH_EINVAL;
    return (int)cudaEventRecord(r->ev_ready[slot], r->s_data);
}
// caller's stream: wait for the slot's K1 (the backbone may read the slot's input from here on)
extern "C" int lh_stage_ring_begin_loss(lh_stage_ring* r, int slot, lh_stream_t stream) {
    if (!r) return LH_ENULL;
    if (slot < 0 || slot >= kSlots) return LH_EINVAL;
    return (int)cudaStreamWaitEvent((cudaStream_t)stream, r->ev_ready[slot], 0);
}
extern "C" int lh_stage_ring_end_loss(lh_stage_ring* r, int slot, lh_stream_t stream) {
    if (!r) return LH_ENULL;
    if (slot < 0 || slot >= kSlots) return LH_EINVAL;
    const cudaError_t e = cudaEventRecord(r->ev_free[slot], (cudaStream_t)stream);
    if (e == cudaSuccess) r->free_valid[slot] = true;
    return (int)e;
}

// the two stages as ONE call each (resident inputs: nothing to wait for but the slot's previous consumer)
extern "C" int lh_stage_ring_data(lh_stage_ring* r, int slot, const lh_warp_args* a) {
    if (!r || !a) return LH_ENULL;
    int rc = lh_stage_ring_begin_data(r, slot, nullptr, nullptr, 0);
    if (rc) return rc;
    rc = lh_warp_normalize(a->src, a->src_img_stride, a->src_h, a->src_w, a->src_pitch, a->fwd_affine, a->flip, a->chan_gain,
                           a->has_mean_std ? a->mean_std : nullptr, a->lut, a->interp, a->chan_order, a->out_layout, a->dst,
                           a->N, a->out_h, a->out_w, r->s_data);
    if (rc) return rc;
    return lh_stage_ring_end_data(r, slot);
}
extern "C" int lh_stage_ring_loss(lh_stage_ring* r, int slot, const lh_heatmap_args* a, lh_stream_t stream) {
    if (!r || !a) return LH_ENULL;
    int rc = lh_stage_ring_begin_loss(r, slot, stream);
    if (rc) return rc;
    rc = lh_heatmap_fused(a, stream);
    if (rc) return rc;
    return lh_stage_ring_end_loss(r, slot, stream);
}

// pinned host allocations for callers that want full-rate DMA into the pipeline
extern "C" void* lh_host_alloc(size_t bytes) {
    void* q = nullptr;
    if (cudaMallocHost(&q, bytes) != cudaSuccess) return nullptr;
    return q;
}
extern "C" void lh_host_free(void* q) { if (q) cudaFreeHost(q); }
