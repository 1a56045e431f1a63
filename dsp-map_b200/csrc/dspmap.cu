// dspmap.cu — host side of the B200 DSP map: handle, HBM layout, per-frame launch sequence, C-ABI.
// Mirrors the public surface of class DSPMap (g-ch/DSP-map include/dsp_dynamic.h:142-446, 1549-1584); the
// per-frame work is done by the sm_100a kernels in dspmap_frame.cuh.  There is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <random>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dspmap_b200.h"
#include "dspmap_frame.cuh"
#include "dspmap_estimator.cuh"
#include "velocity_estimator.h"
#include "host_worker.h"
#include "sparse_rows.h"

namespace {

thread_local std::string g_err;

enum Family {
    FAM_SETUP = 0, FAM_OBS, FAM_ENUM, FAM_PREDICT, FAM_ARRIVE, FAM_PYRAMID, FAM_CK, FAM_WEIGHT, FAM_NORM, FAM_NEWBORN,
    FAM_RESAMPLE, FAM_CLEANUP, FAM_READER, FAM_MISC, FAM_COLL, FAM_EST, FAM_COUNT
};
const char *kFamilyNames[FAM_COUNT] = {"setup", "obs_bin", "enumerate", "predict", "arrive", "pyramid_lists", "ck_pass",
                                       "weight_pass", "newborn_norm", "newborn", "resample_future", "cleanup", "reader", "misc",
                                       "collectives", "velocity_estimation"};

struct ProfSlot {
    cudaEvent_t a, b;
    int fam;
    const char *kernel;
};

}  // namespace

struct dspmap_shard;  // shard_host.inc

struct dspmap {
    dspmap_config cfg;
    dspmap_shard *shard = nullptr;  // part of a sharded map orchestrated by the library (dspmap_shard_init / _init_local)
    MapConst mc;
    DevPtrs dp;
    cudaStream_t stream = nullptr, own_stream = nullptr, side = nullptr, nb = nullptr;  // frame; observation binning + normaliser; early newborn placement
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_state = nullptr, ev_fork_obs = nullptr, ev_join_obs = nullptr;
    cudaEvent_t ev_arrived = nullptr, ev_nb_early = nullptr, ev_staged_t = nullptr;
    bool state_event_recorded = false;
    std::vector<void *> allocs;
    // host mirrors
    std::vector<float> ptab, vtab, lut, planes0;
    std::vector<int> nbr;
    float p_std = 0.2f, v_std = 0.1f, sigma_ob = 0.2f, kappa = 0.01f, Pd = 0.95f;  // :154-158
    float nb_weight = 0.04f;                                                        // :162
    int nb_num = 20;                                                                // :163
    bool tables_dirty = true;
    bool sigma_dirty = true;  // the fast division by sigma must be re-verified
    int fast_sigma = 0;
    int *d_bad = nullptr;
    bool have_last = false;
    float last_p[3] = {0, 0, 0};
    double last_t = 0;
    bool nb_latched = false;
    int nb_min_static = 0, nb_model_gen = 0;
    float update_time = 0.f;
    int update_counter = 0;
    int record_flag = 0;
    float record_time = 1.f;
    bool recorded_once = false;
    std::string record_prefix = "./";  // the CSV's path up to "particles_update_t_..." (the headers differ: dyn:333 adds "/", mn:335 and st:330 do not)
    int stage_limit = 4;
    int max_points = 0, cap_cand = 0;
    // pinned staging
    float *h_pts = nullptr, *h_tagged = nullptr, *h_future = nullptr, *h_xyz = nullptr;
    DevState *h_state = nullptr;
    int *h_count = nullptr;
    int n_tagged = 0;               // size of the current newborn input
    std::vector<float> tagged_host; // last newborn input (world frame), for getKMClusterResult
    bool tagged_registered = false; // tagged_host's reserved storage is page-locked (cudaHostRegister): copied from in place
    void *pinned_user = nullptr;  // caller buffer registered with dspmap_pin_host_buffer
    size_t pinned_bytes = 0;
    // reader scratch
    int *d_blockcnt = nullptr, *d_blockoff = nullptr, *d_count = nullptr;
    float *d_xyz = nullptr, *d_future = nullptr;
    int occ_blocks = 0;
    int occ_guess = 4096;  // occupied voxels copied along with the count (the last count + 25 %): one round trip, not two
    long long upd_h2d_bytes = 0, upd_d2h_bytes = 0;  // what the last dspmap_update call moved over PCIe (cloud, newborn input or cluster velocities; cluster features)
    long long last_d2h_bytes = 0;  // what the last blocking reader call moved over PCIe (count, list, future grid or its rows)
    // sparse copy-out of the future grid into a registered (page-locked) caller buffer (DSPMAP_SPARSE_FUTURE=1)
    int *d_fidx = nullptr, *d_nf = nullptr, *h_fidx = nullptr, *h_nf = nullptr;
    float *d_fval = nullptr, *h_fval = nullptr;
    int fut_guess = 8192;                 // rows copied along with their count
    SparseRows sparse_rows;               // which caller buffer holds the previous call's result, and its non-zero rows
    // pipelined reader (dspmap_get_occupancy_async): two result slots, copies on their own stream
    struct ReaderSlot {
        float *d_xyz = nullptr, *d_future = nullptr, *h_xyz = nullptr, *h_future = nullptr;
        int *d_count = nullptr, *h_count = nullptr;
        cudaEvent_t done = nullptr;
        int copied = 0;
        bool with_future = false, pending = false;
    } rslot[2];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_reader = nullptr;
    int next_slot = 0;
    // ordered prediction noise (dsp_dynamic.h:653-659) stays armed while particles with vz != 0 may exist:
    // constructor-seeded particles until their first prediction, or an injected state that contains such particles
    bool vz_mode = false;
    int vz_blocks = 0;
    // the recompute kernels (k_ck / k_weight) are launched only while the pair buffer may overflow
    bool pyr_scan_fused = true;   // the pyramid table fits the scatter kernel's shared memory
    bool fallback_armed = true;
    bool fallback_forced = false;  // a frame overran the pair buffer without the recompute kernels: keep them armed from now on
    int overflow_latched = 0;      // capacity overruns seen on the device and not yet reported to the caller (codes OR-ed)
    bool clear_overflow = false;   // the device-side flag has been absorbed: clear it in front of the next frame
    long long state_copies = 0, state_absorbed = 0;  // frame-end state copies enqueued / taken over by the host
    bool norm_join_pending = false;  // k_norm runs on the side stream and has not been joined yet
    bool nb_early_done = false;   // this frame's early newborn kernels (candidates, placement) are already enqueued on the newborn branch
    bool pdl = true;              // programmatic dependent launch of the frame's kernels (DSPMAP_PDL=0 turns it off)
    bool async_update = true;     // dspmap_update returns once the frame is enqueued; the next call that needs results waits (DSPMAP_ASYNC_UPDATE=0: wait in update)
    bool staged_pending = false;  // the page-locked staging buffers are still being read by the previous frame's copies
    cudaEvent_t ev_staged = nullptr;
    int nb_pos = 0;               // where frame_a enqueues the early newborn kernels (DSPMAP_NB_POS: 0 behind the estimator, 1 before the weight pass, 2 last)
    bool norm_poll = true;        // k_norm beside the C_z pass, waiting for each 1 / C_z (DSPMAP_NORM_POLL=0: behind it)
    bool timeline = false;        // DSPMAP_TIMELINE=1: the frame's events carry timestamps (dspmap_timeline; diagnosis only)
    bool est_thread = true;       // velocity estimation on the helper thread, beside the enqueueing of the frame (DSPMAP_EST_THREAD=0: calling thread)
    HostWorker worker;
    FrameConst shard_fc;  // frame scalars carried across the phases of a sharded frame
    const float *shard_pts = nullptr;  // this frame's cloud (device), for the binning kernels of phase 1
    int shard_cap_g = 0;
    long long host_u_cur = 0;  // uniform draws consumed on the host while seeding
    VelocityEstimator estimator;
    // front half of the velocity estimation on the device (dspmap_estimator.cuh; DSPMAP_EST_GPU=1, default: host implementation)
    bool est_gpu = false;
    EstPtrs est;
    EstConst est_ec;              // this frame's constants (prepare_estimator)
    bool est_cluster = false;
    unsigned est_hash_mask = 0;
    size_t est_hash_bytes = 0;
    int est_n_pad = 0;            // entries of the device tagged cloud that are initialised (real ones + padding)
    int est_nt_explicit = -1;     // >= 0: the tagged cloud on the device was supplied by the caller and has this many entries
    int est_n_real = 0;           // real entries of the device tagged cloud
    bool tagged_host_stale = false;  // tagged_host has to be fetched from the device (getKMClusterResult)
    int *est_h_hdr = nullptr;
    EstFeature *est_h_feat = nullptr;
    float *est_h_cvel = nullptr;
    float4 *est_d_cvel = nullptr;
    cudaEvent_t ev_feat = nullptr;
    // statistics
    long long launches_total = 0, launches_frame = 0;
    DevState last_state;
    bool profile = false;
    std::vector<ProfSlot> prof_slots;
    size_t prof_used = 0;
    double prof_ms[FAM_COUNT] = {0};
    int prof_n[FAM_COUNT] = {0};
    std::map<std::string, std::pair<double, int>> prof_kernel;  // per kernel name: summed ms, launches
};

namespace {

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            g_err = std::string(#x) + ": " + cudaGetErrorString(e_);                            \
            return DSPMAP_E_CUDA;                                                               \
        }                                                                                       \
    } while (0)

// inside dspmap_create: a failure releases the half-built handle (device memory, streams, events) before returning
#define CKM(x)                                                                                  \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            g_err = std::string(#x) + ": " + cudaGetErrorString(e_);                            \
            dspmap_destroy(m);                                                                  \
            return DSPMAP_E_CUDA;                                                               \
        }                                                                                       \
    } while (0)

template <typename T>
int dalloc(dspmap *m, T **p, size_t n, bool zero = true) {
    void *q = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CK(cudaMalloc(&q, bytes));
    if (zero) CK(cudaMemsetAsync(q, 0, bytes, m->stream));
    m->allocs.push_back(q);
    *p = (T *)q;
    return DSPMAP_OK;
}

void prof_begin(dspmap *m, int fam, const char *kernel) {
    if (!m->profile) return;
    if (m->prof_used == m->prof_slots.size()) {
        ProfSlot s;
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
        m->prof_slots.push_back(s);
    }
    m->prof_slots[m->prof_used].fam = fam;
    m->prof_slots[m->prof_used].kernel = kernel;
    cudaEventRecord(m->prof_slots[m->prof_used].a, m->stream);
}
void prof_end(dspmap *m) {
    if (!m->profile) return;
    cudaEventRecord(m->prof_slots[m->prof_used].b, m->stream);
    ++m->prof_used;
}
void prof_collect(dspmap *m) {
    for (size_t i = 0; i < m->prof_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, m->prof_slots[i].a, m->prof_slots[i].b) == cudaSuccess) {
            m->prof_ms[m->prof_slots[i].fam] += ms;
            m->prof_n[m->prof_slots[i].fam] += 1;
            auto &k = m->prof_kernel[m->prof_slots[i].kernel];
            k.first += ms;
            k.second += 1;
        }
    }
    m->prof_used = 0;
}

// The device state copy that ends every frame has landed in h_state: take it over, latch capacity overruns.
void absorb_state(dspmap *m) {
    m->last_state = *m->h_state;
    if (m->state_absorbed == m->state_copies) return;  // this copy has been looked at already
    m->state_absorbed = m->state_copies;
    if (m->last_state.overflow && !m->clear_overflow) {
        m->overflow_latched |= m->last_state.overflow;
        if (m->last_state.overflow & 4) m->fallback_forced = true;
        m->clear_overflow = true;  // the device flag is sticky until the host has seen it
    }
}
// Reports (once) what absorb_state latched.
int report_overflow(dspmap *m) {
    if (!m->overflow_latched) return DSPMAP_OK;
    char buf[256];
    snprintf(buf, sizeof(buf), "device capacity exceeded: code %d (1 live list, 2 newborn candidates, 4 pair buffer without fallback, "
             "8 shard crossers, 16 shard gather, 32 pyramid-list overflow on a sharded map, 64 normaliser never saw the C_z pass); pairs last frame %llu, capacity %lld", m->overflow_latched,
             m->last_state.total_pairs, m->mc.cap_pairs);
    g_err = buf;
    m->overflow_latched = 0;
    return DSPMAP_E_CAPACITY;
}

// One launch site for every kernel of a frame.  With dspmap::pdl the kernel is launched for programmatic dependent
// launch (see pdl_enter in dspmap_kernels.cuh): its CTAs may become resident while the previous kernel of the stream
// drains, which removes the launch gap between the ~30 short, dependent kernels of a frame.
template <typename... KArgs, typename... Args>
inline void launch_kernel(bool pdl, cudaStream_t st, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args &&...args) {
    if (!pdl) {
        kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
        return;
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)grid, 1, 1);
    lc.blockDim = dim3((unsigned)block, 1, 1);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
}

#define LAUNCH(m, fam, kernel, grid, block, smem, ...)                              \
    do {                                                                            \
        prof_begin(m, fam, #kernel);                                                \
        launch_kernel((m)->pdl, (m)->stream, kernel, (grid), (block), (smem), __VA_ARGS__); \
        prof_end(m);                                                                \
        ++(m)->launches_total;                                                      \
        ++(m)->launches_frame;                                                      \
    } while (0)

// the same on an explicit stream (kernels of a forked branch of the frame); profiled on that stream
#define LAUNCH_ON(m, fam, st, kernel, grid, block, smem, ...)                       \
    do {                                                                            \
        cudaStream_t keep_ = (m)->stream;                                           \
        (m)->stream = (st);                                                         \
        prof_begin(m, fam, #kernel);                                                \
        launch_kernel((m)->pdl, (st), kernel, (grid), (block), (smem), __VA_ARGS__); \
        prof_end(m);                                                                \
        (m)->stream = keep_;                                                        \
        ++(m)->launches_total;                                                      \
        ++(m)->launches_frame;                                                      \
    } while (0)

// the library's A/B switches are environment variables read once, by dspmap_create
bool env_off(const char *name) {  // defaults are on; NAME=0 turns one off
    const char *e = getenv(name);
    return e && strcmp(e, "0") == 0;
}

const int kSMs = 148;
// the two configurations of the C_z chain kernel (threads, floats per tile, rows per tile)
const auto k_cz_wide = &k_cz_chain<256, 8192, 128>;
inline int grid_for(long long n, int block, int max_blocks = kSMs * 8) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// generateGaussianRandomsVectorZeroCenter (dsp_dynamic.h:1150-1160); seeded from cfg.table_seed instead of time(NULL)
int gen_tables(dspmap *m) {
    const int G = m->mc.G;
    std::default_random_engine random(m->cfg.table_seed);
    std::normal_distribution<double> n1(0, m->p_std);
    std::normal_distribution<double> n2(0, m->v_std);
    m->ptab.resize(G);
    m->vtab.resize(G);
    for (int i = 0; i < G; i++) {
        m->ptab[i] = n1(random);
        m->vtab[i] = n2(random);
    }
    CK(cudaMemcpyAsync((void *)m->dp.ptab, m->ptab.data(), sizeof(float) * G, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync((void *)m->dp.vtab, m->vtab.data(), sizeof(float) * G, cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->tables_dirty = false;
    return DSPMAP_OK;
}

int upload_particles(dspmap *m, const int32_t *ids, const float *vals, int n) {
    const MapConst &mc = m->mc;
    CK(cudaMemsetAsync(m->dp.M, 0, sizeof(ulonglong2) * (size_t)mc.V, m->stream));
    if (n > 0) {
        int *d_ids;
        float *d_vals;
        CK(cudaMalloc(&d_ids, sizeof(int) * 2 * (size_t)n));
        CK(cudaMalloc(&d_vals, sizeof(float) * 8 * (size_t)n));
        CK(cudaMemcpyAsync(d_ids, ids, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
        CK(cudaMemcpyAsync(d_vals, vals, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
        LAUNCH(m, FAM_MISC, k_load_scatter, grid_for(n, 256), 256, 0, mc, m->dp, d_ids, d_vals, n);
        CK(cudaStreamSynchronize(m->stream));
        cudaFree(d_ids);
        cudaFree(d_vals);
    }
    return DSPMAP_OK;
}

// addRandomParticles (dsp_dynamic.h:594-624) on the host; the store is empty at this point, so "first free slot"
// (addAParticle, :1183-1201) is simply the next slot of the voxel.  Seeded particles carry the newborn flag 15:
// the first frame's prediction skips them (:649) and the second one draws their velocity noise (:653-659).
int seed_particles(dspmap *m) {
    const MapConst &mc = m->mc;
    const int n = m->cfg.init_particle_num;
    if (n <= 0) return DSPMAP_OK;
    std::vector<int> fill(mc.V, 0), ids;
    std::vector<float> vals;
    u64 k = 0;
    const u64 seed = m->cfg.uniform_seed;
    for (int i = 0; i < n; i++) {
        float px = dsp_uniform(seed, k, -mc.hx, mc.hx), py = dsp_uniform(seed, k + 1, -mc.hy, mc.hy), pz = dsp_uniform(seed, k + 2, -mc.hz, mc.hz);
        float vx = dsp_uniform(seed, k + 3, -1.f, 1.f), vy = dsp_uniform(seed, k + 4, -1.f, 1.f), vz = dsp_uniform(seed, k + 5, -1.f, 1.f);
        k += 6;
        int idx = dsp_voxel_index(mc, px, py, pz);
        if (idx < 0) continue;
        if (fill[idx] >= mc.S) continue;
        ids.push_back(idx);
        ids.push_back(fill[idx]++);
        const float rec[8] = {15.f, vx, vy, vz, px, py, pz, m->cfg.init_weight};
        vals.insert(vals.end(), rec, rec + 8);
    }
    m->host_u_cur = (long long)k;
    m->vz_mode = true;
    return upload_particles(m, ids.data(), vals.data(), (int)ids.size() / 2);
}

// Enables dsp_div_known for divisor b only if it equals IEEE division bit for bit on every float in [-amax, amax].
int verify_fast_div(dspmap *m, float b, float amax, int mode, int *ok) {
    *ok = 0;
    if (!(b > 0.f) || !(amax > 0.f) || !std::isfinite(b) || !std::isfinite(amax)) return DSPMAP_OK;
    unsigned max_bits;
    memcpy(&max_bits, &amax, 4);
    const float r = 1.f / b;
    CK(cudaMemsetAsync(m->d_bad, 0, sizeof(int), m->stream));
    k_verify_div<<<kSMs * 16, 256, 0, m->stream>>>(b, r, max_bits, mode, m->d_bad);
    int bad = 1;
    CK(cudaMemcpyAsync(&bad, m->d_bad, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    *ok = bad == 0;
    return DSPMAP_OK;
}

int ensure_cand_capacity(dspmap *m) {
    int need = m->max_points * std::max(m->nb_num, 1);
    if (need <= m->cap_cand) return DSPMAP_OK;
    m->cap_cand = need;
    if (dalloc(m, &m->dp.CA, need, false) != DSPMAP_OK) return DSPMAP_E_CUDA;
    if (dalloc(m, &m->dp.Caddr, need, false) != DSPMAP_OK) return DSPMAP_E_CUDA;
    if (dalloc(m, &m->dp.Ckey, need, false) != DSPMAP_OK) return DSPMAP_E_CUDA;
    if (dalloc(m, &m->dp.Cdst, need, false) != DSPMAP_OK) return DSPMAP_E_CUDA;
    if (dalloc(m, &m->dp.cseg, need, false) != DSPMAP_OK) return DSPMAP_E_CUDA;
    if (dalloc(m, &m->dp.csegi, need, false) != DSPMAP_OK) return DSPMAP_E_CUDA;
    m->dp.cap_cand = need;
    return DSPMAP_OK;
}

// The tagged cloud was laid out on the device (dspmap_estimator.cuh): bring the host copy up to date.
int fetch_tagged(dspmap *m) {
    if (!m->tagged_host_stale) return DSPMAP_OK;
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaStreamSynchronize(m->nb));
    CK(cudaStreamSynchronize(m->stream));
    m->tagged_host.resize((size_t)7 * m->est_n_real);
    if (m->est_n_real > 0) CK(cudaMemcpy(m->tagged_host.data(), m->dp.tagged, sizeof(float) * 7 * (size_t)m->est_n_real, cudaMemcpyDeviceToHost));
    m->tagged_host_stale = false;
    return DSPMAP_OK;
}
// Front half of the velocity estimation (the reference's side thread, dsp_dynamic.h:1377-1447) on the newborn branch: it
// starts with the frame, beside the prediction chain on the main stream; the early newborn kernels follow it on the same
// stream.  fc->n_tagged becomes the padded size of the tagged cloud (dspmap_estimator.cuh).
void prepare_estimator(dspmap *m, FrameConst *fc, int n) {
    EstConst &ec = m->est_ec;
    memset(&ec, 0, sizeof(ec));
    const MapConst &mc = m->mc;
    const float *planes0 = m->planes0.data();
    dsp_rotate(planes0, fc->q, fc->qi, ec.nrm);
    dsp_rotate(planes0 + 3 * mc.Nh, fc->q, fc->qi, ec.nrm + 3);
    dsp_rotate(planes0 + 3 * (mc.Nh + 1), fc->q, fc->qi, ec.nrm + 6);
    dsp_rotate(planes0 + 3 * (mc.Nh + 1 + mc.Nv), fc->q, fc->qi, ec.nrm + 9);
    for (int k = 0; k < 4; ++k) { ec.q[k] = fc->q[k]; ec.qi[k] = fc->qi[k]; }
    for (int k = 0; k < 3; ++k) ec.cur[k] = fc->cur[k];
    ec.filter_res = m->estimator.filter_res;
    const float tol = 2 * m->estimator.filter_res;  // :1411
    const bool cluster = mc.model != 1 && tol > 0.f && n > 0;
    ec.tol2 = tol * tol;
    ec.inv_cell = cluster ? 1.f / (tol * 0.57f) : 1.f;  // cell edge below tolerance / sqrt(3): a cell's points are mutually linked
    ec.n = n;
    ec.model = mc.model;
    ec.hash_mask = m->est_hash_mask;
    ec.n_pad_prev = m->est_nt_explicit >= 0 ? m->est_nt_explicit : m->est_n_pad;
    ec.nt_override = m->est_nt_explicit;
    ec.n_pad = std::max(ec.n_pad_prev, n);
    m->est_nt_explicit = -1;
    m->est_n_pad = ec.n_pad;
    m->est_cluster = cluster;
    fc->n_tagged = ec.n_pad;
    fc->tagged_padded = 1;
}
int enqueue_estimator(dspmap *m, const float *d_pts) {
    const EstConst &ec = m->est_ec;
    const bool cluster = m->est_cluster;
    const int n = ec.n;
    EstPtrs ep = m->est;
    ep.pts = d_pts;
    cudaStream_t nb = m->nb;
    const int B = 256;
    CK(cudaMemsetAsync(ep.hkey, 0xFF, m->est_hash_bytes, nb));
    CK(cudaMemsetAsync(ep.cnt, 0, sizeof(int) * EC_PER_FRAME, nb));
    LAUNCH_ON(m, FAM_EST, nb, k_est_classify, grid_for(n, B), B, 0, ec, ep);
    if (cluster) {
        LAUNCH_ON(m, FAM_EST, nb, k_est_scatter, grid_for(n, B), B, 0, ec, ep);
        LAUNCH_ON(m, FAM_EST, nb, k_est_link, grid_for(62ll * n, B, kSMs * 16), B, 0, ec, ep);
    }
    LAUNCH_ON(m, FAM_EST, nb, k_est_label, grid_for(n, B), B, 0, ec, ep);
    LAUNCH_ON(m, FAM_EST, nb, k_est_clusters, 1, 1024, 0, ec, ep);
    LAUNCH_ON(m, FAM_EST, nb, k_est_features, cluster ? kSMs * 8 : 1, B, 0, ec, ep);
    CK(cudaEventRecord(m->ev_feat, nb));
    LAUNCH_ON(m, FAM_EST, nb, k_est_write, grid_for(ec.n_pad, B), B, 0, ec, ep);
    CK(cudaGetLastError());
    return DSPMAP_OK;
}
// Second half: waits for the cluster centroids, matches them against the previous frame's on the host (Hungarian, :1449-1499)
// and sends the velocities back; k_est_apply runs on the main stream in front of the kernels that read them.
int finish_estimator(dspmap *m, const FrameConst &fc) {
    CK(cudaEventSynchronize(m->ev_feat));
    const int *hdr = m->est_h_hdr;
    m->est_n_real = hdr[EC_NTAGGED];
    m->tagged_host_stale = true;
    if (hdr[EC_NV] == 0) return DSPMAP_OK;  // nothing in view: the previous cloud, clusters and colours stay (:1379)
    const int ndyn = hdr[EC_NDYN];
    m->upd_d2h_bytes = (long long)sizeof(int) * EC_COUNT + (long long)sizeof(EstFeature) * ndyn;  // written by the device into mapped memory
    m->upd_h2d_bytes += (long long)sizeof(float) * 4 * ndyn;
    m->estimator.finish_device(m->est_h_feat, ndyn, hdr[EC_NC], fc.dt, m->est_h_cvel);
    if (ndyn > 0) {
        CK(cudaMemcpyAsync(m->est_d_cvel, m->est_h_cvel, sizeof(float) * 4 * (size_t)ndyn, cudaMemcpyHostToDevice, m->stream));
        LAUNCH(m, FAM_EST, k_est_apply, grid_for(hdr[EC_NDYNPTS], 256), 256, 0, m->est);
    }
    return DSPMAP_OK;
}

// The early half of the newborn step on its own branch (see k_nb_cand): in-map test and position-noise cursors of the points,
// candidate positions, grouping by destination voxel, slot assignment.  Needs the masks after the arrival pass (ev_arrived).
int enqueue_newborn_early(dspmap *m, const FrameConst &fc, const float *d_tagged) {
    if (!(fc.stage_limit >= 3 && fc.n_tagged > 0 && fc.nb_num > 0)) return DSPMAP_OK;
    const MapConst &mc = m->mc;
    DevPtrs dp = m->dp;
    dp.tagged = d_tagged;
    const int B = 256;
    cudaStream_t nb = m->nb;
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_nb_point0, grid_for(fc.n_tagged, B), B, 0, mc, fc, dp);
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.ninmap, dp.nrank, nullptr, 0, fc.n_tagged}, ScanJob{}, ScanJob{}}});
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_nb_mask, grid_for((long long)fc.n_tagged * fc.nb_num, B), B, 0, mc, fc, dp);
    CK(cudaStreamWaitEvent(nb, m->ev_arrived, 0));
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_nb_cand, grid_for((long long)fc.n_tagged * fc.nb_num, B), B, 0, mc, fc, dp);
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_group_owner, kSMs * 2, B, 0, dp, &dp.st->n_cand_owner, dp.cowner, dp.ccnt, dp.cbase, &dp.st->cand_top);
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_group_scatter, kSMs * 4, B, 0, &dp.st->n_cand, dp.Cdst, dp.Ckey, dp.cbase, dp.cfill, dp.cseg, dp.csegi);
    LAUNCH_ON(m, FAM_NEWBORN, nb, k_nb_place, kSMs * 8, B, 0, mc, fc, dp);
    CK(cudaEventRecord(m->ev_nb_early, nb));
    m->nb_early_done = true;
    return DSPMAP_OK;
}

// Enqueue the first half of a frame: binning, prediction, reassignment, pyramid lists, C_z pass, weight pass.
int enqueue_frame_a(dspmap *m, const FrameConst &fc, const float *d_pts, const float *d_tagged_early = nullptr, bool device_estimator = false) {
    const MapConst &mc = m->mc;
    DevPtrs dp = m->dp;
    dp.pts = d_pts;
    m->launches_frame = 0;
    const int B = 256;
    int rc_nb = DSPMAP_OK;
    LAUNCH(m, FAM_SETUP, k_frame_setup, 1, 256, 0, mc, fc, dp);
    CK(cudaEventRecord(m->ev_fork_obs, m->stream));
    // (Order of the calls below: when a frame starts on an idle device — the host-pointer API — kernels run as soon as they are
    // enqueued, so the host feeds the critical chain first: prediction and reassignment, then the two side branches, then the rest.)
    // prediction and reassignment
    if (fc.vz_mode) {
        LAUNCH(m, FAM_PREDICT, k_vz_count, grid_for(mc.V, B), B, 0, mc, dp);
        LAUNCH(m, FAM_PREDICT, k_scan_blocksum, m->vz_blocks, 256, 0, dp.vzcnt, mc.V, dp.vzblk);
        LAUNCH(m, FAM_PREDICT, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.vzblk, dp.vzblkoff, nullptr, 0, m->vz_blocks}, ScanJob{}, ScanJob{}}});
        LAUNCH(m, FAM_PREDICT, k_scan_apply, m->vz_blocks, 256, 0, dp.vzcnt, mc.V, dp.vzblkoff, dp.vzoff, m->vz_blocks);
    }
    LAUNCH(m, FAM_ENUM, k_enumerate, grid_for(mc.V, B), B, 0, mc, dp, 1);
    LAUNCH(m, FAM_PREDICT, k_predict, kSMs * 8, B, 0, mc, fc, dp);
    if (fc.vz_mode) LAUNCH(m, FAM_PREDICT, k_vz_advance, 1, 32, 0, mc, dp);
    LAUNCH(m, FAM_ARRIVE, k_group_owner, kSMs * 2, B, 0, dp, &dp.st->n_mov_owner, dp.mowner, dp.mcnt, dp.mbase, &dp.st->mov_top);
    LAUNCH(m, FAM_ARRIVE, k_group_scatter, kSMs * 4, B, 0, &dp.st->n_mov, dp.MBdst, dp.MBkey, dp.mbase, dp.mfill, dp.mseg, (int *)nullptr);
    LAUNCH(m, FAM_ARRIVE, k_arrive, kSMs * 4, B, 0, mc, fc, dp);
    CK(cudaEventRecord(m->ev_arrived, m->stream));  // the occupancy masks are final until the newborn placement
    // the newborn branch starts behind this frame's setup (i.e. behind everything of the previous frame)
    CK(cudaStreamWaitEvent(m->nb, m->ev_fork_obs, 0));
    m->nb_early_done = false;
    if (device_estimator && (rc_nb = enqueue_estimator(m, d_pts)) != DSPMAP_OK) return rc_nb;  // its tagged cloud feeds the early newborn kernels
    if (m->nb_pos == 0) { if (d_tagged_early && (rc_nb = enqueue_newborn_early(m, fc, d_tagged_early)) != DSPMAP_OK) return rc_nb; }
    // observations: binning touches nothing the prediction / reassignment chain reads, and both are chains of small
    // latency-bound kernels, so they run side by side (joined before the pair preparation, the first consumer of the bins)
    CK(cudaStreamWaitEvent(m->side, m->ev_fork_obs, 0));
    if (fc.n_points > 0) {
        LAUNCH_ON(m, FAM_OBS, m->side, k_obs_classify, grid_for(fc.n_points, B), B, 0, mc, fc, dp);
    }
    LAUNCH_ON(m, FAM_OBS, m->side, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.obs_cnt, dp.obs_off, dp.obs_capoff, mc.OBS - 1, mc.P}, ScanJob{}, ScanJob{}}});
    if (fc.n_points > 0) {
        LAUNCH_ON(m, FAM_OBS, m->side, k_obs_scatter, grid_for(fc.n_points, B), B, 0, mc, fc, dp);
        LAUNCH_ON(m, FAM_OBS, m->side, k_obs_rank, grid_for(fc.n_points, B), B, 0, mc, fc, dp);
    }
    CK(cudaEventRecord(m->ev_join_obs, m->side));
    if (m->pyr_scan_fused) {  // every block of the scatter kernel scans the pyramid counts itself (shared memory)
        LAUNCH(m, FAM_PYRAMID, k_pyr_scatter, kSMs * 2, B, sizeof(int) * (mc.P + 1), mc, dp, 1);
    } else {
        LAUNCH(m, FAM_PYRAMID, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.pcount, dp.poff, nullptr, 0, mc.P}, ScanJob{}, ScanJob{}}});
        LAUNCH(m, FAM_PYRAMID, k_pyr_scatter, kSMs * 8, B, 0, mc, dp, 0);
    }
    LAUNCH(m, FAM_PYRAMID, k_pyr_sort, std::min(mc.P, kSMs * 3), 512, PYR_SORT_CAP * sizeof(u64), mc, dp, fc.Pd);
    CK(cudaStreamWaitEvent(m->stream, m->ev_join_obs, 0));
    if (fc.stage_limit >= 2) {
        LAUNCH(m, FAM_CK, k_pair_prep, mc.P > 2048 ? std::min((mc.P + 1023) / 1024, 32) : 1, 1024, 0, mc, dp);
        if (fc.stage_limit >= 3 && m->norm_poll) {  // the newborn normaliser is one long serial chain: it runs beside the C_z pass and takes each 1 / C_z as it appears
            CK(cudaEventRecord(m->ev_fork, m->stream));
            CK(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
            launch_kernel(m->pdl, m->side, k_norm, 1, 128, 0, mc, fc, dp, 1);
            ++m->launches_total;
            ++m->launches_frame;
            CK(cudaEventRecord(m->ev_join, m->side));
        }
        LAUNCH(m, FAM_CK, k_pair_eval, kSMs * 2, EVAL_THREADS, EVAL_SMEM_BYTES, mc, fc, dp, 0);
        LAUNCH(m, FAM_CK, k_cz_wide, std::min(mc.P, kSMs * 3), 256, sizeof(float) * (2 * (8192 + 8) + 2 * 128), mc, fc, dp);
        size_t smem4 = sizeof(float) * (DSP_LUT_HALF + 3 + K4_TERMS) + sizeof(float4) * (256 + mc.OBS);
        if (m->fallback_armed) LAUNCH(m, FAM_CK, k_ck, std::min(mc.P, kSMs * 2), K4_THREADS, smem4, mc, fc, dp);  // returns at once when the pair buffer is used
        if (fc.stage_limit >= 3 && !m->norm_poll) {  // DSPMAP_NORM_POLL=0: behind the C_z pass, beside the weight pass
            CK(cudaEventRecord(m->ev_fork, m->stream));
            CK(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
            launch_kernel(m->pdl, m->side, k_norm, 1, 128, 0, mc, fc, dp, 0);
            ++m->launches_total;
            ++m->launches_frame;
            CK(cudaEventRecord(m->ev_join, m->side));
        }
        if (m->nb_pos == 1) { if (d_tagged_early && (rc_nb = enqueue_newborn_early(m, fc, d_tagged_early)) != DSPMAP_OK) return rc_nb; }
        LAUNCH(m, FAM_WEIGHT, k_weight2, kSMs * 8, W2_THREADS, 0, mc, fc, dp);
        LAUNCH(m, FAM_WEIGHT, k_weight2w, kSMs * 6, W2W_THREADS, 0, mc, fc, dp);
        size_t smem5 = sizeof(float) * (DSP_LUT_HALF + 3) + sizeof(float4) * (size_t)mc.NB * (mc.OBS - 1);
        int chunks = (mc.L + K5_THREADS - 1) / K5_THREADS;
        if (m->fallback_armed) LAUNCH(m, FAM_WEIGHT, k_weight, kSMs * 2, K5_THREADS, smem5, mc, fc, dp, chunks);
        // the normaliser is first read by k_nb_cand (w_new): it is joined behind the newborn kernels that do not need it
        m->norm_join_pending = fc.stage_limit >= 3;
    }
    // With a device-resident newborn input the early newborn kernels (they need the cloud, the noise table and the masks after
    // the arrival pass) run beside the observation passes; with a host cloud they are enqueued by enqueue_frame_b, once the cloud is there
    if (m->nb_pos >= 2 || (m->nb_pos == 1 && fc.stage_limit < 2)) { if (d_tagged_early && (rc_nb = enqueue_newborn_early(m, fc, d_tagged_early)) != DSPMAP_OK) return rc_nb; }
    CK(cudaGetLastError());
    return DSPMAP_OK;
}
// Second half: newborn particles (needs the velocity-tagged cloud), occupancy + resampling + future status.
int enqueue_frame_b(dspmap *m, const FrameConst &fc, const float *d_tagged) {
    const MapConst &mc = m->mc;
    DevPtrs dp = m->dp;
    dp.tagged = d_tagged;
    const int B = 256;
    int newborn_ran = 0;
    if (fc.stage_limit >= 3 && fc.n_tagged > 0 && fc.nb_num > 0) {
        int rc;
        if (!m->nb_early_done && (rc = enqueue_newborn_early(m, fc, d_tagged)) != DSPMAP_OK) return rc;
        CK(cudaStreamWaitEvent(m->stream, m->ev_nb_early, 0));  // the newborn slots exist (flag 15): the split below skips them
        LAUNCH(m, FAM_NEWBORN, k_nb_point1, grid_for((long long)fc.n_tagged * 32, B), B, 0, mc, fc, dp, 0);
        LAUNCH(m, FAM_NEWBORN, k_scan_small, 2, 1024, 0, ScanJobs{{ScanJob{dp.nvcnt, dp.nvoff, nullptr, 0, fc.n_tagged}, ScanJob{dp.nrcnt, dp.nroff, nullptr, 0, fc.n_tagged}, ScanJob{}}});
        if (m->norm_join_pending) {  // k_norm (side stream) wrote w_new, which k_nb_fill reads
            CK(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
            m->norm_join_pending = false;
        }
        LAUNCH(m, FAM_NEWBORN, k_nb_fill, kSMs * 8, B, 0, mc, fc, dp, (u64)m->cfg.uniform_seed);
        newborn_ran = 1;
    }
    if (fc.stage_limit >= 4) {
        LAUNCH(m, FAM_RESAMPLE, k_voxel_list, grid_for(mc.V, B), B, 0, mc, dp);
        LAUNCH(m, FAM_RESAMPLE, k_resample, kSMs * (RS_VPW >= 4 ? 4 : 16), 32 * RS_WARPS, RS_WARPS * rs_warp_bytes(mc.S), mc, fc, dp);
    }
    if (m->norm_join_pending) {  // no newborn kernels this frame: k_norm must still be over before the next frame resets its outputs
        CK(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
        m->norm_join_pending = false;
    }
    LAUNCH(m, FAM_CLEANUP, k_cleanup, kSMs * 2, B, 0, mc, fc, dp, newborn_ran, m->fallback_armed ? 1 : 0);
    // the state copy steers which optional kernels the next frame launches
    CK(cudaMemcpyAsync(m->h_state, m->dp.st, sizeof(DevState), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaEventRecord(m->ev_state, m->stream));
    m->state_event_recorded = true;
    ++m->state_copies;
    CK(cudaGetLastError());
    return DSPMAP_OK;
}

// DSPMap::update's host prologue (dsp_dynamic.h:186-224, 292-293): validation, odometry delta, frame scalars
int frame_prologue(dspmap *m, int n, float px, float py, float pz, double t, float qw, float qx, float qy, float qz, FrameConst *fc) {
    if (!m->have_last) {  // function-local statics initialised by the first call (:187-190)
        m->last_p[0] = px; m->last_p[1] = py; m->last_p[2] = pz;
        m->last_t = t;
        m->have_last = true;
    }
    if (std::fabs(qw) > 1.001f || std::fabs(qx) > 1.001f || std::fabs(qy) > 1.001f || std::fabs(qz) > 1.001f) {
        printf("Invalid quaternion.\n");
        return DSPMAP_REJECTED;
    }
    float ox = px - m->last_p[0], oy = py - m->last_p[1], oz = pz - m->last_p[2];
    float dt = (float)(t - m->last_t);
    if (std::fabs(ox) > 10.f || std::fabs(oy) > 10.f || std::fabs(oz) > 10.f || dt < 0.f || dt > 10.f) {
        printf("!!! delt_t = %f\n", dt);
        return DSPMAP_REJECTED;
    }
    m->last_p[0] = px; m->last_p[1] = py; m->last_p[2] = pz;
    m->last_t = t;
    m->update_time += dt;  // :634-635
    m->update_counter += 1;
    if (!m->nb_latched) {  // function-local statics of the newborn step, frozen at first use (:808-811; st:791)
        m->nb_min_static = (int)((float)m->nb_num * (m->mc.model == 1 ? 0.2f : 0.15f));
        m->nb_model_gen = (int)((float)m->nb_num * 0.8f);
        m->nb_latched = true;
    }
    memset(fc, 0, sizeof(*fc));
    fc->q[0] = qw; fc->q[1] = qx; fc->q[2] = qy; fc->q[3] = qz;
    dsp_quat_inverse(fc->q, fc->qi);
    fc->sx = -ox; fc->sy = -oy; fc->sz = -oz;
    fc->dt = dt;
    fc->cur[0] = px; fc->cur[1] = py; fc->cur[2] = pz;
    fc->sigma = m->sigma_ob;
    fc->sigma_r = 1.f / m->sigma_ob;
    fc->fast_sigma = m->fast_sigma;
    fc->Pd = m->Pd;
    fc->one_minus_Pd = 1 - m->Pd;
    fc->kappa = m->kappa;
    fc->nb_weight = m->nb_weight;
    fc->nb_num = m->nb_num;
    fc->nb_min_static = m->nb_min_static;
    fc->nb_model_gen = m->nb_model_gen;
    fc->n_points = n;
    fc->stage_limit = m->stage_limit;
    fc->vz_mode = m->vz_mode ? 1 : 0;
    // The recompute kernels are left out only when the host KNOWS the previous frame's pair count (its state copy has
    // completed) and it is below half the buffer; growth from one frame to the next is gradual.
    m->fallback_armed = true;
    if (m->state_event_recorded && cudaEventQuery(m->ev_state) == cudaSuccess) {
        absorb_state(m);
        if (m->update_counter > 2) m->fallback_armed = m->fallback_forced || m->h_state->total_pairs * 2ull > (unsigned long long)m->mc.cap_pairs;
    } else {
        cudaGetLastError();  // cudaErrorNotReady is not an error here
    }
    if (m->clear_overflow) {  // stream-ordered in front of this frame's kernels
        if (cudaMemsetAsync(&m->dp.st->overflow, 0, sizeof(int), m->stream) != cudaSuccess) { g_err = "cudaMemsetAsync(overflow)"; return DSPMAP_E_CUDA; }
        m->clear_overflow = false;
    }
    return DSPMAP_OK;
}

int frame_epilogue(dspmap *m) {
    CK(cudaStreamSynchronize(m->stream));  // (every frame ends with an asynchronous copy of the device state)
    absorb_state(m);
    if (m->profile) prof_collect(m);
    // once a frame has predicted every particle and drew no noise, all vz are 0 for good (LIMIT_MOVEMENT_IN_XY_PLANE)
    if (m->vz_mode && m->stage_limit >= 4 && m->last_state.n_vz == 0 && m->last_state.n_skipped == 0) m->vz_mode = false;
    return report_overflow(m);
}

int write_particle_csv(dspmap *m);

// pyramid boundary-plane normals when the sensor has no rotation (dsp_dynamic.h:563-578)
void make_planes0(const dspmap_config *cfg, int Nh, int Nv, std::vector<float> &planes0) {
    // :543 `(float)angle_resolution / 180.f * M_PIf32`: glibc's <cmath> defines M_PIf32 as a FLOAT literal under _GNU_SOURCE
    // (g++'s default), so the header's own double fallback (:77-79) is not used and the product is rounded in fp32
    const float ang = cfg->pi_is_double ? (float)((float)cfg->angle_resolution / 180.f * 3.14159265358979323846)
                                        : (float)cfg->angle_resolution / 180.f * 3.14159265358979323846f;
    planes0.assign(3 * (Nh + Nv + 2), 0.f);
    int h0 = -cfg->half_fov_h / cfg->angle_resolution, h1 = -h0;
    for (int i = h0; i <= h1; i++) {
        planes0[3 * (i + h1) + 0] = -std::sin((float)i * ang);
        planes0[3 * (i + h1) + 1] = std::cos((float)i * ang);
        planes0[3 * (i + h1) + 2] = 0.f;
    }
    float *pv = &planes0[3 * (Nh + 1)];
    int v0 = -cfg->half_fov_v / cfg->angle_resolution, v1 = -v0;
    for (int i = v0; i <= v1; i++) {
        pv[3 * (i + v1) + 0] = std::sin((float)i * ang);
        pv[3 * (i + v1) + 1] = 0.f;
        pv[3 * (i + v1) + 2] = std::cos((float)i * ang);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char *dspmap_last_error(void) { return g_err.c_str(); }

void dspmap_default_config(dspmap_config *c) {
    memset(c, 0, sizeof(*c));
    c->nx = 66; c->ny = 66; c->nz = 40;
    c->resolution = 0.15f;
    c->angle_resolution = 3;
    c->half_fov_h = 42; c->half_fov_v = 24;
    c->max_particles_per_voxel = 9;
    c->pyramid_neighbor_n = 1;
    c->model = 0;
    c->prediction_times = 6;
    const float ft[6] = {0.05f, 0.2f, 0.5f, 1.f, 1.5f, 2.f};
    memcpy(c->prediction_future_time, ft, sizeof(ft));
    c->occlusion_margin = 0.3f;
    c->init_particle_num = 0;
    c->init_weight = 0.01f;
    // the reference seeds both generators from the wall clock (dsp_dynamic.h:586, 1151); DSPMAP_TABLE_SEED / DSPMAP_UNIFORM_SEED
    // pin them for reproducible runs of an unchanged application (tests/test_dropin.py compares such a run bit for bit)
    c->table_seed = (uint64_t)time(nullptr);
    c->uniform_seed = (uint64_t)time(nullptr);
    if (const char *e = getenv("DSPMAP_TABLE_SEED")) if (*e) c->table_seed = strtoull(e, nullptr, 10);
    if (const char *e = getenv("DSPMAP_UNIFORM_SEED")) if (*e) c->uniform_seed = strtoull(e, nullptr, 10);
    c->gaussian_table_size = 10000000;
    c->max_observations_per_pyramid = 100;
    c->device = 0;
    c->max_points = 65536;
}

int dspmap_create(const dspmap_config *cfg, dspmap **out) {
    if (!cfg || !out) { g_err = "null argument"; return DSPMAP_E_BAD_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_err = "no CUDA device: this library has no CPU path";
        return DSPMAP_E_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major < 10) {
        g_err = "device is not sm_100 class: the kernels are built for sm_100a only";
        return DSPMAP_E_NO_DEVICE;
    }
    if (cfg->nx <= 0 || cfg->ny <= 0 || cfg->nz <= 0 || !(cfg->resolution > 0.f) || cfg->angle_resolution <= 0 || cfg->half_fov_h <= 0 ||
        cfg->half_fov_v <= 0 || cfg->max_particles_per_voxel <= 0 || cfg->pyramid_neighbor_n < 0 || cfg->prediction_times < 0 ||
        (long long)cfg->nx * cfg->ny * cfg->nz * DSP_MAX_SLOTS > 2147483647ll) {
        g_err = "bad configuration: sizes, resolutions and field-of-view angles must be positive and V * 128 < 2^31";
        return DSPMAP_E_BAD_ARG;
    }
    CK(cudaSetDevice(cfg->device));
    dspmap *m = new dspmap();
    m->cfg = *cfg;
    MapConst &mc = m->mc;
    memset(&mc, 0, sizeof(mc));
    mc.nx = cfg->nx; mc.ny = cfg->ny; mc.nz = cfg->nz;
    mc.V = cfg->nx * cfg->ny * cfg->nz;
    mc.res = cfg->resolution;
    mc.hx = (mc.res * (float)mc.nx) * 0.5f;  // :528-530
    mc.hy = (mc.res * (float)mc.ny) * 0.5f;
    mc.hz = (mc.res * (float)mc.nz) * 0.5f;
    mc.Nh = cfg->half_fov_h * 2 / cfg->angle_resolution;  // :58-60
    mc.Nv = cfg->half_fov_v * 2 / cfg->angle_resolution;
    mc.P = mc.Nh * mc.Nv;
    const int pyramid_num = 360 * 180 / cfg->angle_resolution / cfg->angle_resolution;  // :63
    const int safe_particle_num = mc.V * cfg->max_particles_per_voxel + 1e5;           // :64
    mc.max_ppv = cfg->max_particles_per_voxel;
    mc.model = cfg->model;
    mc.S = cfg->safe_particles_per_voxel > 0 ? cfg->safe_particles_per_voxel : mc.max_ppv * (mc.model == 1 ? 5 : 2);  // :65, st:63
    mc.L = cfg->safe_particles_per_pyramid > 0 ? cfg->safe_particles_per_pyramid : safe_particle_num / pyramid_num * 2;  // :66
    mc.T = cfg->prediction_times;
    mc.NB = (2 * cfg->pyramid_neighbor_n + 1) * (2 * cfg->pyramid_neighbor_n + 1);
    mc.NBW = mc.NB + 1;
    mc.OBS = cfg->max_observations_per_pyramid > 0 ? cfg->max_observations_per_pyramid : 100;
    mc.G = cfg->gaussian_table_size > 0 ? cfg->gaussian_table_size : 10000000;
    mc.occl = cfg->occlusion_margin;
    mc.z_begin = 0;
    mc.z_end = mc.nz;
    mc.sharded = 0; mc.rank = 0; mc.nranks = 1; mc.z_per_rank = mc.nz; mc.v_lo = 0; mc.v_hi = mc.V;
    for (int i = 0; i < DSP_MAX_T; ++i) mc.ft[i] = cfg->prediction_future_time[i];
    if (mc.S > DSP_MAX_SLOTS || mc.S < 1 || mc.T > DSP_MAX_T || mc.T < 0 || mc.V <= 0 || mc.P <= 0 ||
        mc.Nh + mc.Nv + 2 > DSP_MAX_PLANES || (long long)mc.V * DSP_MAX_SLOTS > 2147483647ll || mc.L < 1 ||
        mc.OBS > 128 || mc.NB * (mc.OBS - 1) * 16 > 150000) {
        g_err = "configuration outside supported limits (S <= 128 slots, T <= 8, V*128 < 2^31)";
        delete m;
        return DSPMAP_E_BAD_ARG;
    }
    mc.vlo = mc.S >= 64 ? ~0ull : ((1ull << mc.S) - 1ull);
    mc.vhi = mc.S <= 64 ? 0ull : (mc.S >= 128 ? ~0ull : ((1ull << (mc.S - 64)) - 1ull));
    m->max_points = cfg->max_points > 0 ? cfg->max_points : 65536;
    // (Stream priorities were tried — main chain high, newborn branch low — and starved the newborn branch: its early kernels
    // ended 0.15 ms later and with them the frame; profiles/r02_variants.jsonl.)
    CKM(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    CKM(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
    CKM(cudaStreamCreateWithFlags(&m->nb, cudaStreamNonBlocking));
    { const char *tl = getenv("DSPMAP_TIMELINE"); m->timeline = tl && strcmp(tl, "1") == 0; }
    const unsigned evflags = m->timeline ? cudaEventDefault : cudaEventDisableTiming;
    CKM(cudaEventCreateWithFlags(&m->ev_arrived, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_nb_early, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_staged_t, cudaEventDisableTiming));
    CKM(cudaEventCreateWithFlags(&m->ev_fork, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_join, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_fork_obs, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_join_obs, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_state, evflags));
    CKM(cudaEventCreateWithFlags(&m->ev_staged, cudaEventDisableTiming));
    m->stream = m->own_stream;

    DevPtrs &dp = m->dp;
    memset(&dp, 0, sizeof(dp));
    const size_t V = mc.V, VS = V * mc.S, P = mc.P, MP = m->max_points;
    long long cap_live = std::min<long long>((long long)V * mc.S, 8ll << 20);
    dp.cap_live = (int)cap_live;
    const size_t CL = (size_t)cap_live;
    int rc = DSPMAP_OK;
#define A(ptr, n) if (rc == DSPMAP_OK) rc = dalloc(m, &ptr, (n))
    A(dp.PA, VS); A(dp.PB, VS); A(dp.M, V); A(dp.M0, V); A(dp.MS, V); A(dp.OCCV, V); A(dp.FUT, V * std::max(mc.T, 1));
    A(dp.E, CL);
    m->vz_blocks = (int)((V + SCAN_BLOCK - 1) / SCAN_BLOCK);
    A(dp.vzcnt, V); A(dp.vzoff, V + 1); A(dp.vzblk, m->vz_blocks + 1); A(dp.vzblkoff, m->vz_blocks + 1);
    A(dp.OR, MP); A(dp.OPID, MP); A(dp.obs_cnt, P); A(dp.obs_fill, P); A(dp.obs_maxbits, P); A(dp.obs_off, P + 1);
    A(dp.obs_capoff, P + 1); A(dp.OSEG, MP); A(dp.OBSP, P * mc.OBS); A(dp.CZ, P * mc.OBS); A(dp.INV, MP + P * 0 + 1024);
    A(dp.MBA, CL); A(dp.MBB, CL); A(dp.MBkey, CL); A(dp.MBdst, CL); A(dp.MBq, CL);
    A(dp.mcnt, V); A(dp.mfill, V); A(dp.mbase, V); A(dp.mowner, V); A(dp.mseg, CL);
    A(dp.Fkey, CL); A(dp.Faddr, CL); A(dp.Fq, CL); A(dp.FP, CL); A(dp.PSpay, CL); A(dp.pcount, P); A(dp.pfill, P); A(dp.pub, P); A(dp.rkey, CL); A(dp.rkey2, CL); A(dp.rpos, CL); A(dp.rcount, V + 1); A(dp.poff, P + 1); A(dp.plen, P);
    A(dp.PSkey, CL); A(dp.PSaddr, CL); A(dp.LA, CL); A(dp.LP, CL); A(dp.PW, CL);
    mc.cap_pairs = 512ll << 20;  // 2 GB of fp32 pair terms (of 180 GB); larger frames fall back to the recompute kernels
    A(dp.G, (size_t)mc.cap_pairs + 64); A(dp.cum, P * mc.NBW); A(dp.totlen, P); A(dp.pairs, P + 1); A(dp.rowbase, P + 1);
    A(dp.chunks, P + 1); A(dp.chunk_off, P + 1); A(dp.chunk_pyr, CL / 32 + P + 1);
    A(dp.NPC, MP); A(dp.ninmap, MP + 1); A(dp.nrank, MP + 1); A(dp.nstatic, MP); A(dp.nvcnt, MP + 1); A(dp.nrcnt, MP + 1);
    A(dp.nvoff, MP + 1); A(dp.nroff, MP + 1); A(dp.nimask, MP);
    A(dp.ccnt, V); A(dp.cfill, V); A(dp.cbase, V); A(dp.cowner, V);
    A(dp.st, 1); A(m->d_bad, 1);
    float *d_ptab, *d_vtab, *d_lut, *d_planes0, *d_pts, *d_tagged;
    int *d_nbr;
    A(d_ptab, mc.G); A(d_vtab, mc.G); A(d_lut, DSP_LUT_HALF); A(d_planes0, 3 * (mc.Nh + mc.Nv + 2)); A(dp.planes, 3 * (mc.Nh + mc.Nv + 2));
    int *d_nbrev;
    A(d_nbr, P * mc.NBW); A(d_nbrev, P * mc.NBW); A(d_pts, MP * 3); A(d_tagged, MP * 7);
    m->occ_blocks = (int)((V + OCC_BLOCK - 1) / OCC_BLOCK);
    A(m->d_blockcnt, m->occ_blocks + 1); A(m->d_blockoff, m->occ_blocks + 1); A(m->d_count, 1); A(m->d_xyz, V * 3);
    A(m->d_future, V * std::max(mc.T, 1));
#undef A
    if (rc != DSPMAP_OK) { dspmap_destroy(m); return rc; }
    dp.ptab = d_ptab; dp.vtab = d_vtab; dp.lut = d_lut; dp.planes0 = d_planes0; dp.nbr = d_nbr; dp.nbrev = d_nbrev;
    dp.pts = d_pts; dp.tagged = d_tagged;
    if (ensure_cand_capacity(m) != DSPMAP_OK) { dspmap_destroy(m); return DSPMAP_E_CUDA; }
    CKM(cudaMallocHost(&m->h_pts, sizeof(float) * MP * 3));
    CKM(cudaMallocHost(&m->h_tagged, sizeof(float) * MP * 7));
    m->tagged_host.reserve((size_t)MP * 7 + 16);  // never reallocated afterwards: every writer stays within max_points entries
    m->tagged_registered = cudaHostRegister(m->tagged_host.data(), sizeof(float) * ((size_t)MP * 7 + 16), cudaHostRegisterDefault) == cudaSuccess;
    if (!m->tagged_registered) cudaGetLastError();
    CKM(cudaMallocHost(&m->h_future, sizeof(float) * V * std::max(mc.T, 1)));
    CKM(cudaMallocHost(&m->h_xyz, sizeof(float) * V * 3));
    CKM(cudaMallocHost(&m->h_state, sizeof(DevState)));
    CKM(cudaMallocHost(&m->h_count, sizeof(int)));
    {   // measured on B200 (profiles/r02_variants.jsonl): end to end the two are level at cfg2 (host 0.565 ms, device 0.565 - 0.58), the
        // host estimator is ahead with the pipelined reader and at cfg3 (the device is left to the frame) — so it is the default
        const char *e = getenv("DSPMAP_EST_GPU");
        m->est_gpu = e && strcmp(e, "1") == 0;
    }
    m->norm_poll = !env_off("DSPMAP_NORM_POLL");
    { const char *e = getenv("DSPMAP_NB_POS"); if (e && e[0] >= '0' && e[0] <= '2') m->nb_pos = e[0] - '0'; }
    if (m->est_gpu) {
        int rc2 = DSPMAP_OK;
        EstPtrs &e = m->est;
        memset(&e, 0, sizeof(e));
        const size_t NP = (size_t)MP + 1, NCL = (size_t)MP / EST_MIN_CLUSTER + 2;
        size_t H = 1024;
        while (H < 4 * (size_t)MP) H <<= 1;
        m->est_hash_mask = (unsigned)(H - 1);
        m->est_hash_bytes = H * (sizeof(u64) + sizeof(int));
        unsigned char *hash = nullptr;
#define AE(ptr, count) if (rc2 == DSPMAP_OK) rc2 = dalloc(m, &ptr, (size_t)(count))
        AE(e.W, NP); AE(e.parent, NP); AE(e.csize, NP); AE(e.label, NP); AE(e.pos, NP); AE(e.grank, NP); AE(e.mrank, NP);
        AE(e.kidx, NP); AE(e.flag_g, NP); AE(e.flag_r, NP); AE(e.cells, NP); AE(e.cbase, NP); AE(e.ccells, NP); AE(e.bbox, 6 * NP); AE(e.SW, NP); AE(e.cellof, NP); AE(e.cmin, NP); AE(e.ccount, NP); AE(e.rootc, NP); AE(e.cellid, H); AE(hash, m->est_hash_bytes);
        AE(e.croot, NCL); AE(e.csz, NCL); AE(e.spos, NCL); AE(e.cdyn, NCL); AE(e.coff, NCL); AE(e.dseq, NCL); AE(e.order, NCL);
        AE(e.a_dyn, NCL); AE(e.a_dsz, NCL); AE(e.a_ssz, NCL); AE(e.p_dyn, NCL); AE(e.p_dsz, NCL); AE(e.p_ssz, NCL); AE(e.cfeat, NCL);
        AE(e.cnt, EC_COUNT); AE(e.tcid, NP); AE(m->est_d_cvel, NCL);
#undef AE
        if (rc2 != DSPMAP_OK) { dspmap_destroy(m); return rc2; }
        e.hkey = (u64 *)hash;
        e.hcnt = (int *)(hash + H * sizeof(u64));
        e.cvel = m->est_d_cvel;
        e.tagged = d_tagged;
        CKM(cudaHostAlloc(&m->est_h_hdr, sizeof(int) * EC_COUNT, cudaHostAllocMapped));
        CKM(cudaHostAlloc(&m->est_h_feat, sizeof(EstFeature) * NCL, cudaHostAllocMapped));
        CKM(cudaMallocHost(&m->est_h_cvel, sizeof(float) * 4 * NCL));
        CKM(cudaHostGetDevicePointer((void **)&e.h_hdr, m->est_h_hdr, 0));
        CKM(cudaHostGetDevicePointer((void **)&e.h_feat, m->est_h_feat, 0));
        CKM(cudaEventCreateWithFlags(&m->ev_feat, evflags));
    }

    make_planes0(cfg, mc.Nh, mc.Nv, m->planes0);  // boundary-plane normals in the sensor frame (:563-578)
    // neighbour table (:1128-1147; mn:1135-1136)
    m->nbr.assign(P * mc.NBW, 0);
    for (int p = 0; p < mc.P; p++) {
        int h = p / mc.Nv, v = p % mc.Nv, n = 0;
        for (int i = -cfg->pyramid_neighbor_n; i <= cfg->pyramid_neighbor_n; ++i)
            for (int j = -cfg->pyramid_neighbor_n; j <= cfg->pyramid_neighbor_n; ++j) {
                int hh = h + i, vv = v + j;
                if (hh >= 0 && hh < mc.Nh && vv >= 0 && vv < mc.Nv) m->nbr[(size_t)p * mc.NBW + 1 + n++] = hh * mc.Nv + vv;
            }
        m->nbr[(size_t)p * mc.NBW] = n;
    }
    // where each pyramid sits in its neighbours' lists (the relation is symmetric)
    std::vector<int> nbrev(P * mc.NBW, 0);
    for (int a = 0; a < mc.P; a++)
        for (int ns = 0; ns < m->nbr[(size_t)a * mc.NBW]; ++ns) {
            const int i = m->nbr[(size_t)a * mc.NBW + 1 + ns];
            for (int k = 0; k < m->nbr[(size_t)i * mc.NBW]; ++k)
                if (m->nbr[(size_t)i * mc.NBW + 1 + k] == a) nbrev[(size_t)a * mc.NBW + 1 + ns] = k;
        }
    // PDF table (:1282-1292), host libm exactly as the reference evaluates it; only indices 10000..20000 are kept:
    // the table is symmetric (x = (i-10000)*0.001f negates exactly and powf(x,2) is even), checked here.
    m->lut.resize(DSP_LUT_HALF);
    {
        std::vector<float> full(20000);
        for (int i = 0; i < 20000; ++i) {
            float x = (float)(i - 10000) * 0.001f;
            full[i] = (1.f / (sqrtf(2.f * 1.57079632679489661923))) * expf(-powf(x, 2) / (2));
        }
        for (int k = 1; k < 10000; ++k)
            if (memcmp(&full[10000 + k], &full[10000 - k], 4) != 0) {
                g_err = "PDF table is not symmetric on this libm";
                dspmap_destroy(m);
                return DSPMAP_E_BAD_ARG;
            }
        for (int h = 0; h < 10000; ++h) m->lut[h] = full[10000 + h];
        m->lut[10000] = full[0];  // |i - 10000| = 10000 only for i = 0 (x = -10); unreachable, queries clamp to |x| <= 9.9
    }
    CKM(cudaMemcpyAsync(d_lut, m->lut.data(), sizeof(float) * DSP_LUT_HALF, cudaMemcpyHostToDevice, m->stream));
    CKM(cudaMemcpyAsync(d_planes0, m->planes0.data(), sizeof(float) * m->planes0.size(), cudaMemcpyHostToDevice, m->stream));
    CKM(cudaMemcpyAsync(d_nbr, m->nbr.data(), sizeof(int) * m->nbr.size(), cudaMemcpyHostToDevice, m->stream));
    CKM(cudaMemcpyAsync(d_nbrev, nbrev.data(), sizeof(int) * nbrev.size(), cudaMemcpyHostToDevice, m->stream));
    CKM(cudaStreamSynchronize(m->stream));  // nbrev is a local
    CKM(cudaFuncSetAttribute(k_pyr_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PYR_SORT_CAP * sizeof(u64))));
    m->pyr_scan_fused = sizeof(int) * (size_t)(mc.P + 1) <= 200 * 1024;
    if (m->pyr_scan_fused) CKM(cudaFuncSetAttribute(k_pyr_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * (mc.P + 1))));
    CKM(cudaFuncSetAttribute(k_ck, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CKM(cudaFuncSetAttribute(k_pair_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    CKM(cudaFuncSetAttribute(k_cz_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    // the three defaults that can be turned off for A/B measurements (profiles/r02_ab_switches.jsonl)
    m->pdl = !env_off("DSPMAP_PDL");
    m->est_thread = !env_off("DSPMAP_EST_THREAD");
    m->async_update = !env_off("DSPMAP_ASYNC_UPDATE");
    CKM(cudaFuncSetAttribute(k_weight, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKM(cudaFuncSetAttribute(k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RS_WARPS * rs_warp_bytes(DSP_MAX_SLOTS))));
    CKM(cudaStreamSynchronize(m->stream));
    if (gen_tables(m) != DSPMAP_OK) { dspmap_destroy(m); return DSPMAP_E_CUDA; }
    {   // (p + half) / res with p inside the map: dividends lie in (0, 2*half)
        mc.res_r = 1.f / mc.res;
        int ok = 0;
        if (verify_fast_div(m, mc.res, 2.f * std::max(mc.hx, std::max(mc.hy, mc.hz)) * 1.0001f, 0, &ok) != DSPMAP_OK) { dspmap_destroy(m); return DSPMAP_E_CUDA; }
        mc.fast_res = ok;
    }
    m->estimator.reset(cfg->uniform_seed);
    rc = seed_particles(m);
    if (rc != DSPMAP_OK) { dspmap_destroy(m); return rc; }
    if (m->host_u_cur) {
        DevState st;
        memset(&st, 0, sizeof(st));
        st.u_cur = m->host_u_cur;
        CKM(cudaMemcpy(dp.st, &st, sizeof(st), cudaMemcpyHostToDevice));
    }
    memset(&m->last_state, 0, sizeof(m->last_state));
    printf("Map is ready to update!\n");  // :174
    *out = m;
    return DSPMAP_OK;
}

void dspmap_shard_release(dspmap *m);

void dspmap_destroy(dspmap *m) {
    if (!m) return;
    m->worker.stop();
    cudaSetDevice(m->cfg.device);
    cudaDeviceSynchronize();
    dspmap_shard_release(m);
    if (m->pinned_user) cudaHostUnregister(m->pinned_user);
    for (void *p : m->allocs) cudaFree(p);
    if (m->h_pts) cudaFreeHost(m->h_pts);
    if (m->tagged_registered) cudaHostUnregister(m->tagged_host.data());
    if (m->h_tagged) cudaFreeHost(m->h_tagged);
    if (m->h_future) cudaFreeHost(m->h_future);
    if (m->h_xyz) cudaFreeHost(m->h_xyz);
    if (m->h_state) cudaFreeHost(m->h_state);
    if (m->h_count) cudaFreeHost(m->h_count);
    if (m->h_fidx) cudaFreeHost(m->h_fidx);
    if (m->h_fval) cudaFreeHost(m->h_fval);
    if (m->h_nf) cudaFreeHost(m->h_nf);
    for (auto &s : m->prof_slots) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->ev_fork_obs) cudaEventDestroy(m->ev_fork_obs);
    if (m->ev_join_obs) cudaEventDestroy(m->ev_join_obs);
    if (m->ev_state) cudaEventDestroy(m->ev_state);
    if (m->ev_staged) cudaEventDestroy(m->ev_staged);
    for (auto &s : m->rslot) {
        if (s.done) cudaEventDestroy(s.done);
        if (s.h_xyz) cudaFreeHost(s.h_xyz);
        if (s.h_future) cudaFreeHost(s.h_future);
        if (s.h_count) cudaFreeHost(s.h_count);
    }
    if (m->ev_reader) cudaEventDestroy(m->ev_reader);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    if (m->side) cudaStreamDestroy(m->side);
    if (m->nb) cudaStreamDestroy(m->nb);
    if (m->ev_arrived) cudaEventDestroy(m->ev_arrived);
    if (m->ev_nb_early) cudaEventDestroy(m->ev_nb_early);
    if (m->ev_staged_t) cudaEventDestroy(m->ev_staged_t);
    if (m->est_h_hdr) cudaFreeHost(m->est_h_hdr);
    if (m->est_h_feat) cudaFreeHost(m->est_h_feat);
    if (m->est_h_cvel) cudaFreeHost(m->est_h_cvel);
    if (m->ev_feat) cudaEventDestroy(m->ev_feat);
    delete m;
}

static int update_common(dspmap *m, int n, int stride, const float *pts, float px, float py, float pz, double t, float qw,
                         float qx, float qy, float qz, const float *tagged, int n_tagged, bool use_estimator) {
    if (!m || n < 0 || stride < 3 || (n > 0 && !pts) || n_tagged < 0) { g_err = "bad argument"; return DSPMAP_E_BAD_ARG; }
    if (n > m->max_points || n_tagged > m->max_points) { g_err = "more points than dspmap_config.max_points"; return DSPMAP_E_CAPACITY; }
    CK(cudaSetDevice(m->cfg.device));
    FrameConst fc;
    int rc = frame_prologue(m, n, px, py, pz, t, qw, qx, qy, qz, &fc);
    if (rc != DSPMAP_OK) return rc;
    if (m->tables_dirty && (rc = gen_tables(m)) != DSPMAP_OK) return rc;
    if (m->sigma_dirty) {  // queries are clamped to |x - mu| / sigma <= 9.9: beyond 16 sigma both divisions clamp alike
        if ((rc = verify_fast_div(m, m->sigma_ob, 16.f * m->sigma_ob, 1, &m->fast_sigma)) != DSPMAP_OK) return rc;
        m->sigma_dirty = false;
        fc.fast_sigma = m->fast_sigma;
    }
    if ((rc = ensure_cand_capacity(m)) != DSPMAP_OK) return rc;
    if (m->staged_pending) {  // asynchronous updates: the previous frame's host-to-device copies must have left the staging buffers
        CK(cudaEventSynchronize(m->ev_staged));
        CK(cudaEventSynchronize(m->ev_staged_t));
        m->staged_pending = false;
    }
    for (int i = 0; i < n; ++i) {
        m->h_pts[3 * i] = pts[(size_t)i * stride];
        m->h_pts[3 * i + 1] = pts[(size_t)i * stride + 1];
        m->h_pts[3 * i + 2] = pts[(size_t)i * stride + 2];
    }
    if (n > 0) CK(cudaMemcpyAsync((void *)m->dp.pts, m->h_pts, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    m->upd_h2d_bytes = (long long)sizeof(float) * 3 * n;
    m->upd_d2h_bytes = 0;
    if (m->async_update) CK(cudaEventRecord(m->ev_staged, m->stream));  // behind the staging copy of the cloud
    const bool on_device = use_estimator && m->est_gpu;
    const bool on_helper = use_estimator && !on_device && m->est_thread;
    if (on_device) prepare_estimator(m, &fc, n);
    if (on_helper) {  // the estimation starts now, on the helper thread, while this thread enqueues the frame (host_worker.h)
        m->worker.start();
        dspmap *mm = m;
        m->worker.submit([mm, fc, n] { mm->estimator.estimate(mm->mc, fc, mm->planes0.data(), mm->h_pts, n, mm->cfg.model, mm->tagged_host); });
    }
    rc = enqueue_frame_a(m, fc, m->dp.pts, on_device ? m->dp.tagged : nullptr, on_device);
    if (on_helper) m->worker.wait();  // joined before the newborn step, like the reference's thread (:311)
    if (rc != DSPMAP_OK) return rc;
    if (on_device) {
        if (m->async_update) {
            CK(cudaEventRecord(m->ev_staged_t, m->nb));
            m->staged_pending = true;
        }
        if ((rc = finish_estimator(m, fc)) != DSPMAP_OK) return rc;
        if ((rc = enqueue_frame_b(m, fc, m->dp.tagged)) != DSPMAP_OK) return rc;
    } else {
    if (!use_estimator && (rc = fetch_tagged(m)) != DSPMAP_OK) return rc;  // (frames of the device estimator came before)
    if (use_estimator && !on_helper) {
        // the reference's side thread (dsp_dynamic.h:297, 1377-1544), overlapped with the kernels enqueued above exactly
        // as the reference overlaps it with prediction + update (:297-311)
        m->estimator.estimate(m->mc, fc, m->planes0.data(), m->h_pts, n, m->cfg.model, m->tagged_host);
    } else if (!use_estimator && tagged) {
        // explicit newborn input.  A NULL pointer means "unchanged": the reference's input_cloud_with_velocity is a member that
        // keeps its previous content when the side thread finds nothing in view (dsp_dynamic.h:1387-1391); a non-NULL pointer
        // with n_tagged == 0 empties it
        m->tagged_host.assign(tagged, tagged + (size_t)7 * n_tagged);
    }
    int nt = (int)(m->tagged_host.size() / 7);
    if (nt > m->max_points) { g_err = "newborn input larger than max_points"; return DSPMAP_E_CAPACITY; }
    if (nt > 0) {
        // on the newborn branch (which is behind the previous frame by now): the copy and the early newborn kernels that
        // follow it there do not queue up behind this frame's observation passes
        // (tagged_host is page-locked in place — its capacity is reserved and registered at create time — so the estimator's
        // output goes to the device without a staging copy: 280 KB less to move on the calling thread, between the join and the newborn kernels)
        const float *src = m->tagged_host.data();
        if (!m->tagged_registered) {
            memcpy(m->h_tagged, m->tagged_host.data(), sizeof(float) * 7 * (size_t)nt);
            src = m->h_tagged;
        }
        CK(cudaMemcpyAsync((void *)m->dp.tagged, src, sizeof(float) * 7 * (size_t)nt, cudaMemcpyHostToDevice, m->nb));
        m->upd_h2d_bytes += (long long)sizeof(float) * 7 * nt;
    }
    fc.n_tagged = nt;
    m->est_nt_explicit = nt;  // (a later frame of the device estimator starts from this cloud)
    m->est_n_real = nt;
    m->tagged_host_stale = false;
    if (m->async_update) {
        CK(cudaEventRecord(m->ev_staged_t, m->nb));  // behind the staging copy of the newborn input
        m->staged_pending = true;
    }
    if ((rc = enqueue_frame_b(m, fc, m->dp.tagged)) != DSPMAP_OK) return rc;
    }
    // Asynchronous update: like dspmap_update_device, return with the frame enqueued; readers and dumps are stream-ordered or
    // synchronise themselves, dspmap_counters / dspmap_synchronize pick up the frame's state copy.  A capacity overrun of an
    // earlier frame (latched by frame_prologue from that frame's state copy) is reported here; this frame's own by the next
    // call that synchronises.
    if (m->async_update && !m->vz_mode && !m->record_flag && !m->profile) return report_overflow(m);
    if ((rc = frame_epilogue(m)) != DSPMAP_OK) return rc;
    // particle CSV (:325-350)
    if (m->record_flag) {
        if (m->record_flag < 0 || (m->update_time > m->record_time && !m->recorded_once)) {
            m->recorded_once = true;
            write_particle_csv(m);
        }
    }
    return DSPMAP_OK;
}

int dspmap_update(dspmap *m, int n, int stride, const float *pts, float px, float py, float pz, double t, float qw,
                  float qx, float qy, float qz) {
    return update_common(m, n, stride, pts, px, py, pz, t, qw, qx, qy, qz, nullptr, 0, true);
}
int dspmap_update_tagged(dspmap *m, int n, int stride, const float *pts, float px, float py, float pz, double t,
                         float qw, float qx, float qy, float qz, const float *tagged, int n_tagged) {
    return update_common(m, n, stride, pts, px, py, pz, t, qw, qx, qy, qz, tagged, n_tagged, false);
}
int dspmap_update_device(dspmap *m, int n, const float *d_pts, float px, float py, float pz, double t, float qw,
                         float qx, float qy, float qz, const float *d_tagged, int n_tagged) {
    if (!m || n < 0 || n > m->max_points || n_tagged < 0 || n_tagged > m->max_points || (n > 0 && !d_pts) || (n_tagged > 0 && !d_tagged)) {
        g_err = "bad argument";
        return DSPMAP_E_BAD_ARG;
    }
    CK(cudaSetDevice(m->cfg.device));
    FrameConst fc;
    int rc = frame_prologue(m, n, px, py, pz, t, qw, qx, qy, qz, &fc);
    if (rc != DSPMAP_OK) return rc;
    if (m->tables_dirty && (rc = gen_tables(m)) != DSPMAP_OK) return rc;
    if (m->sigma_dirty) {  // queries are clamped to |x - mu| / sigma <= 9.9: beyond 16 sigma both divisions clamp alike
        if ((rc = verify_fast_div(m, m->sigma_ob, 16.f * m->sigma_ob, 1, &m->fast_sigma)) != DSPMAP_OK) return rc;
        m->sigma_dirty = false;
        fc.fast_sigma = m->fast_sigma;
    }
    if ((rc = ensure_cand_capacity(m)) != DSPMAP_OK) return rc;
    fc.n_tagged = n_tagged;
    if ((rc = enqueue_frame_a(m, fc, d_pts, d_tagged)) != DSPMAP_OK) return rc;
    if ((rc = enqueue_frame_b(m, fc, d_tagged)) != DSPMAP_OK) return rc;
    if (m->vz_mode) return frame_epilogue(m);  // keep the ordered-noise path armed only as long as it is needed
    return report_overflow(m);  // an earlier frame's overrun, latched by frame_prologue; this frame's own shows up at the next synchronising call
}

int dspmap_shard_config(dspmap *m, int rank, int nranks, float *xsend, float *xrecv, int cap_x, float *gsend, float *grecv,
                        int cap_g, float *czinv, float *shared) {
    if (!m || nranks < 1 || rank < 0 || rank >= nranks || cap_x < 1 || cap_g < 1 || !xsend || !xrecv || !gsend || !grecv || !czinv || !shared) {
        g_err = "bad shard configuration";
        return DSPMAP_E_BAD_ARG;
    }
    MapConst &mc = m->mc;
    if ((long long)nranks * cap_g > m->dp.cap_live) { g_err = "nranks * cap_g exceeds the live-particle capacity"; return DSPMAP_E_BAD_ARG; }
    mc.sharded = 1;
    mc.rank = rank;
    mc.nranks = nranks;
    mc.z_per_rank = (mc.nz + nranks - 1) / nranks;
    const int z0 = std::min(mc.nz, rank * mc.z_per_rank), z1 = std::min(mc.nz, (rank + 1) * mc.z_per_rank);
    mc.z_begin = z0;
    mc.z_end = z1;
    mc.v_lo = z0 * mc.nx * mc.ny;
    mc.v_hi = z1 * mc.nx * mc.ny;
    mc.cap_x = cap_x;
    mc.cap_g = cap_g;
    m->shard_cap_g = cap_g;
    DevPtrs &dp = m->dp;
    dp.xsend = xsend; dp.xrecv = xrecv; dp.gsend = gsend; dp.grecv = grecv;
    dp.CZ = czinv;                                   // C_z and 1/C_z live in the caller's buffer: merged by all-reduce
    dp.INV = czinv + (size_t)mc.P * mc.OBS;
    dp.nst_shared = shared;                          // newborn split first, then the new weights by global list index
    dp.NW = shared + m->max_points;
    m->fallback_armed = false;  // sharded frames always use the pair buffer (checked: overflow flag otherwise)
    return DSPMAP_OK;
}
int dspmap_shard_gather_records(dspmap *m, int records) {
    if (!m || !m->mc.sharded || records < 1 || records > m->shard_cap_g) { g_err = "bad gather size"; return DSPMAP_E_BAD_ARG; }
    m->mc.cap_g = records;
    return DSPMAP_OK;
}

int dspmap_shard_phase(dspmap *m, int phase, int n, const float *d_pts, float px, float py, float pz, double t, float qw,
                       float qx, float qy, float qz, const float *d_tagged, int n_tagged) {
    if (!m || !m->mc.sharded) { g_err = "handle is not sharded"; return DSPMAP_E_BAD_ARG; }
    CK(cudaSetDevice(m->cfg.device));
    const int B = 256;
    int rc;
    if (phase == 0) {
        if (n < 0 || n > m->max_points || n_tagged > m->max_points) { g_err = "bad argument"; return DSPMAP_E_BAD_ARG; }
        m->mc.cap_g = m->shard_cap_g;  // the pack kernel may fill the whole slab; the frame's stride is set after phase 1
        FrameConst fc;
        rc = frame_prologue(m, n, px, py, pz, t, qw, qx, qy, qz, &fc);
        if (rc != DSPMAP_OK) return rc;
        if (m->tables_dirty && (rc = gen_tables(m)) != DSPMAP_OK) return rc;
        if (m->sigma_dirty) {
            if ((rc = verify_fast_div(m, m->sigma_ob, 16.f * m->sigma_ob, 1, &m->fast_sigma)) != DSPMAP_OK) return rc;
            m->sigma_dirty = false;
            fc.fast_sigma = m->fast_sigma;
        }
        if ((rc = ensure_cand_capacity(m)) != DSPMAP_OK) return rc;
        fc.n_tagged = n_tagged;
        m->fallback_armed = false;
        m->shard_fc = fc;
    }
    const MapConst &mc = m->mc;
    const FrameConst &fc = m->shard_fc;
    DevPtrs dp = m->dp;
    if (phase == 0) {
        dp.pts = d_pts;
        m->launches_frame = 0;
        m->shard_pts = d_pts;
        LAUNCH(m, FAM_SETUP, k_frame_setup, 1, 256, 0, mc, fc, dp);
        // observation binning (replicated on every rank) on the side branch, beside prediction, exchange and arrival; joined in phase 2
        CK(cudaEventRecord(m->ev_fork_obs, m->stream));
        CK(cudaStreamWaitEvent(m->side, m->ev_fork_obs, 0));
        if (fc.n_points > 0) LAUNCH_ON(m, FAM_OBS, m->side, k_obs_classify, grid_for(fc.n_points, B), B, 0, mc, fc, dp);
        LAUNCH_ON(m, FAM_OBS, m->side, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.obs_cnt, dp.obs_off, dp.obs_capoff, mc.OBS - 1, mc.P}, ScanJob{}, ScanJob{}}});
        if (fc.n_points > 0) {
            LAUNCH_ON(m, FAM_OBS, m->side, k_obs_scatter, grid_for(fc.n_points, B), B, 0, mc, fc, dp);
            LAUNCH_ON(m, FAM_OBS, m->side, k_obs_rank, grid_for(fc.n_points, B), B, 0, mc, fc, dp);
        }
        CK(cudaEventRecord(m->ev_join_obs, m->side));
        CK(cudaStreamWaitEvent(m->nb, m->ev_fork_obs, 0));  // the newborn branch starts behind this frame's setup
        m->nb_early_done = false;
        if (fc.vz_mode) {
            LAUNCH(m, FAM_PREDICT, k_vz_count, grid_for(mc.V, B), B, 0, mc, dp);
            LAUNCH(m, FAM_PREDICT, k_scan_blocksum, m->vz_blocks, 256, 0, dp.vzcnt, mc.V, dp.vzblk);
            LAUNCH(m, FAM_PREDICT, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.vzblk, dp.vzblkoff, nullptr, 0, m->vz_blocks}, ScanJob{}, ScanJob{}}});
            LAUNCH(m, FAM_PREDICT, k_scan_apply, m->vz_blocks, 256, 0, dp.vzcnt, mc.V, dp.vzblkoff, dp.vzoff, m->vz_blocks);
        }
        LAUNCH(m, FAM_ENUM, k_enumerate, grid_for(mc.V, B), B, 0, mc, dp, 1);
        LAUNCH(m, FAM_PREDICT, k_predict, kSMs * 8, B, 0, mc, fc, dp);
        if (fc.vz_mode) LAUNCH(m, FAM_PREDICT, k_vz_advance, 1, 32, 0, mc, dp);
        LAUNCH(m, FAM_PREDICT, k_shard_headers, 1, 32, 0, mc, dp);  // the slab headers now bound every rank's registered particles
    } else if (phase == 1) {
        dp.pts = m->shard_pts;
        LAUNCH(m, FAM_ARRIVE, k_shard_import, grid_for((long long)mc.nranks * mc.cap_x, B), B, 0, mc, dp);
        LAUNCH(m, FAM_ARRIVE, k_group_owner, kSMs * 2, B, 0, dp, &dp.st->n_mov_owner, dp.mowner, dp.mcnt, dp.mbase, &dp.st->mov_top);
        LAUNCH(m, FAM_ARRIVE, k_group_scatter, kSMs * 4, B, 0, &dp.st->n_mov, dp.MBdst, dp.MBkey, dp.mbase, dp.mfill, dp.mseg, (int *)nullptr);
        LAUNCH(m, FAM_ARRIVE, k_arrive, kSMs * 4, B, 0, mc, fc, dp);
        CK(cudaEventRecord(m->ev_arrived, m->stream));  // the occupancy masks are final until the newborn placement
        LAUNCH(m, FAM_PYRAMID, k_shard_pack_fov, kSMs * 4, B, 0, mc, dp);
        // WHERE this rank's newborn particles land depends only on the cloud, the noise table and the masks after the arrival pass:
        // candidates, grouping and placement run on the newborn branch, beside the gather and the observation passes (like on one GPU)
        if ((rc = enqueue_newborn_early(m, fc, d_tagged)) != DSPMAP_OK) return rc;
    } else if (phase == 2) {
        LAUNCH(m, FAM_PYRAMID, k_shard_fov_gathered, kSMs * 8, B, 0, mc, dp, 0);
        LAUNCH(m, FAM_PYRAMID, k_scan_small, 1, 1024, 0, ScanJobs{{ScanJob{dp.pcount, dp.poff, nullptr, 0, mc.P}, ScanJob{}, ScanJob{}}});
        LAUNCH(m, FAM_PYRAMID, k_shard_fov_gathered, kSMs * 8, B, 0, mc, dp, 1);
        LAUNCH(m, FAM_PYRAMID, k_pyr_sort, std::min(mc.P, kSMs * 3), 512, PYR_SORT_CAP * sizeof(u64), mc, dp, fc.Pd);
        CK(cudaStreamWaitEvent(m->stream, m->ev_join_obs, 0));
        LAUNCH(m, FAM_CK, k_pair_prep, mc.P > 2048 ? std::min((mc.P + 1023) / 1024, 32) : 1, 1024, 0, mc, dp);
        LAUNCH(m, FAM_CK, k_shard_zero, kSMs * 2, B, 0, mc, fc, dp, 0);
        LAUNCH(m, FAM_CK, k_pair_eval, kSMs * 2, EVAL_THREADS, EVAL_SMEM_BYTES, mc, fc, dp, 1);
        LAUNCH(m, FAM_CK, k_cz_wide, std::min(mc.P, kSMs * 3), 256, sizeof(float) * (2 * (8192 + 8) + 2 * 128), mc, fc, dp);
    } else if (phase == 3) {
        dp.tagged = d_tagged;
        if (fc.stage_limit >= 3) {  // 1 / C_z is complete on every rank (merged behind phase 2): the normaliser's serial chain runs beside the weight pass
            CK(cudaEventRecord(m->ev_fork, m->stream));
            CK(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
            LAUNCH_ON(m, FAM_NORM, m->side, k_norm, 1, 128, 0, mc, fc, dp, 0);
            CK(cudaEventRecord(m->ev_join, m->side));
            m->norm_join_pending = true;
        }
        LAUNCH(m, FAM_WEIGHT, k_shard_zero, kSMs * 2, B, 0, mc, fc, dp, 1);
        LAUNCH(m, FAM_WEIGHT, k_pair_eval, kSMs * 2, EVAL_THREADS, EVAL_SMEM_BYTES, mc, fc, dp, 2);
        LAUNCH(m, FAM_WEIGHT, k_weight2, kSMs * 8, W2_THREADS, 0, mc, fc, dp);
        LAUNCH(m, FAM_WEIGHT, k_weight2w, kSMs * 6, W2W_THREADS, 0, mc, fc, dp);
    } else if (phase == 4) {  // owners take their new weights; the newborn split reads them (dsp_dynamic.h:829-866)
        dp.tagged = d_tagged;
        LAUNCH(m, FAM_WEIGHT, k_shard_apply_weights, kSMs * 4, B, 0, mc, dp);
        if (fc.n_tagged > 0 && fc.nb_num > 0) {
            LAUNCH(m, FAM_NEWBORN, k_shard_zero, kSMs, B, 0, mc, fc, dp, 2);
            if (!m->nb_early_done && (rc = enqueue_newborn_early(m, fc, d_tagged)) != DSPMAP_OK) return rc;  // (a driver that skipped phase 1's early half)
            CK(cudaStreamWaitEvent(m->stream, m->ev_nb_early, 0));  // point pass 0 ran there; the slots born early carry flag 15: the split skips them
            LAUNCH(m, FAM_NEWBORN, k_nb_point1, grid_for((long long)fc.n_tagged * 32, B), B, 0, mc, fc, dp, 1);
        }
    } else if (phase == 5) {
        dp.tagged = d_tagged;
        int newborn_ran = 0;
        if (fc.n_tagged > 0 && fc.nb_num > 0) {
            LAUNCH(m, FAM_NEWBORN, k_nb_point1, grid_for((long long)fc.n_tagged * 32, B), B, 0, mc, fc, dp, 2);
            LAUNCH(m, FAM_NEWBORN, k_scan_small, 2, 1024, 0, ScanJobs{{ScanJob{dp.nvcnt, dp.nvoff, nullptr, 0, fc.n_tagged}, ScanJob{dp.nrcnt, dp.nroff, nullptr, 0, fc.n_tagged}, ScanJob{}}});
            if (m->norm_join_pending) {  // k_norm (side branch) wrote w_new, which k_nb_fill reads
                CK(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
                m->norm_join_pending = false;
            }
            LAUNCH(m, FAM_NEWBORN, k_nb_fill, kSMs * 8, B, 0, mc, fc, dp, (u64)m->cfg.uniform_seed);
            newborn_ran = 1;
        }
        if (m->norm_join_pending) {
            CK(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
            m->norm_join_pending = false;
        }
        LAUNCH(m, FAM_RESAMPLE, k_voxel_list, grid_for(mc.V, B), B, 0, mc, dp);
        LAUNCH(m, FAM_RESAMPLE, k_resample, kSMs * (RS_VPW >= 4 ? 4 : 16), 32 * RS_WARPS, RS_WARPS * rs_warp_bytes(mc.S), mc, fc, dp);
        LAUNCH(m, FAM_CLEANUP, k_cleanup, kSMs * 2, B, 0, mc, fc, dp, newborn_ran, 0);
        CK(cudaMemcpyAsync(m->h_state, m->dp.st, sizeof(DevState), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaEventRecord(m->ev_state, m->stream));
        m->state_event_recorded = true;
        ++m->state_copies;
    } else {
        g_err = "phase must be 0..5";
        return DSPMAP_E_BAD_ARG;
    }
    CK(cudaGetLastError());
    return DSPMAP_OK;
}

int dspmap_set_prediction_variance(dspmap *m, float p, float v) {
    if (!m) return DSPMAP_E_BAD_ARG;
    m->p_std = p;
    m->v_std = v;
    m->tables_dirty = true;  // regenerated before the next frame (cursors are kept, :355-360)
    return DSPMAP_OK;
}
int dspmap_set_observation_stddev(dspmap *m, float s) {
    if (!m) return DSPMAP_E_BAD_ARG;
    m->sigma_ob = s;
    m->sigma_dirty = true;
    return DSPMAP_OK;
}
int dspmap_set_newborn_weight(dspmap *m, float w) { if (!m) return DSPMAP_E_BAD_ARG; m->nb_weight = w; return DSPMAP_OK; }
int dspmap_set_newborn_number(dspmap *m, int n) {
    if (!m || n < 0 || n > DSP_MAX_NB_NUM) { g_err = "newborn number must be in [0, 64]"; return DSPMAP_E_BAD_ARG; }
    m->nb_num = n;
    return DSPMAP_OK;
}
int dspmap_set_particle_record_flag(dspmap *m, int flag, float record_time, const char *prefix) {
    if (!m) return DSPMAP_E_BAD_ARG;
    m->record_flag = flag;
    m->record_time = record_time;
    if (prefix) m->record_prefix = prefix;
    return DSPMAP_OK;
}
int dspmap_set_voxel_filter_resolution(dspmap *m, float r) { if (!m) return DSPMAP_E_BAD_ARG; m->estimator.filter_res = r; return DSPMAP_OK; }

int dspmap_get_occupancy_device(dspmap *m, float thr, float *d_xyz, int cap, int *d_count, float *d_future) {
    if (!m) return DSPMAP_E_BAD_ARG;
    CK(cudaSetDevice(m->cfg.device));
    const MapConst &mc = m->mc;
    LAUNCH(m, FAM_READER, k_occ_count, m->occ_blocks, 256, 0, mc, m->dp, thr, m->d_blockcnt, d_future, (int *)nullptr, (float *)nullptr, (int *)nullptr);
    LAUNCH(m, FAM_READER, k_occ_write, m->occ_blocks, 256, 0, mc, m->dp, thr, m->d_blockcnt, d_xyz, cap, d_count, m->occ_blocks);
    CK(cudaGetLastError());
    return DSPMAP_OK;
}
int dspmap_get_occupancy(dspmap *m, float thr, float *xyz_out, int cap, int *n_out, float *future) {
    if (!m) return DSPMAP_E_BAD_ARG;
    CK(cudaSetDevice(m->cfg.device));
    const MapConst &mc = m->mc;
    const size_t fbytes = sizeof(float) * (size_t)mc.V * mc.T;
    const bool direct = future && m->pinned_user && (char *)future >= (char *)m->pinned_user &&
                        (char *)future + fbytes <= (char *)m->pinned_user + m->pinned_bytes;
    // A registered buffer is only ever written by this function, so it still holds the previous call's result: instead of
    // the dense grid, the non-zero voxel rows travel and the buffer is patched (rows that were non-zero last time are cleared
    // first).  Contract (include/dspmap_b200.h): while registered, the application treats the buffer as read-only.
    const bool sparse = direct && mc.T > 0;
    int fguess = 0;
    if (sparse) {
        if (!m->d_fidx) {
            if (dalloc(m, &m->d_fidx, (size_t)mc.V, false) != DSPMAP_OK || dalloc(m, &m->d_fval, (size_t)mc.V * mc.T, false) != DSPMAP_OK ||
                dalloc(m, &m->d_nf, 2) != DSPMAP_OK)  // d_nf[0]: rows, d_nf[1]: occupied voxels (one copy brings both)
                return DSPMAP_E_CUDA;
            CK(cudaMallocHost(&m->h_fidx, sizeof(int) * (size_t)mc.V));
            CK(cudaMallocHost(&m->h_fval, sizeof(float) * (size_t)mc.V * mc.T));
            CK(cudaMallocHost(&m->h_nf, 2 * sizeof(int)));
        }
        // the reader's first kernel packs the non-zero rows while it reads and clears the grid (no dense device copy, no extra kernels)
        CK(cudaMemsetAsync(m->d_nf, 0, sizeof(int), m->stream));
        LAUNCH(m, FAM_READER, k_occ_count, m->occ_blocks, 256, 0, mc, m->dp, thr, m->d_blockcnt, (float *)nullptr, m->d_fidx, m->d_fval, m->d_nf);
        LAUNCH(m, FAM_READER, k_occ_write, m->occ_blocks, 256, 0, mc, m->dp, thr, m->d_blockcnt, m->d_xyz, mc.V, m->d_nf + 1, m->occ_blocks);
        CK(cudaGetLastError());
    } else {
        int rc = dspmap_get_occupancy_device(m, thr, m->d_xyz, mc.V, m->d_count, future ? m->d_future : nullptr);
        if (rc != DSPMAP_OK) return rc;
    }
    if (!sparse) CK(cudaMemcpyAsync(m->h_count, m->d_count, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    // the occupied-voxel list travels with its count: as many records as the last frames suggest, the rest (rare) after
    const int guess = xyz_out ? std::min(std::min(m->occ_guess, cap), mc.V) : 0;
    if (guess > 0) CK(cudaMemcpyAsync(m->h_xyz, m->d_xyz, sizeof(float) * 3 * (size_t)guess, cudaMemcpyDeviceToHost, m->stream));
    if (sparse) {
        fguess = std::min(m->fut_guess, mc.V);
        CK(cudaMemcpyAsync(m->h_nf, m->d_nf, 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaMemcpyAsync(m->h_fidx, m->d_fidx, sizeof(int) * (size_t)fguess, cudaMemcpyDeviceToHost, m->stream));
        CK(cudaMemcpyAsync(m->h_fval, m->d_fval, sizeof(float) * (size_t)fguess * mc.T, cudaMemcpyDeviceToHost, m->stream));
    } else if (future) {
        CK(cudaMemcpyAsync(direct ? future : m->h_future, m->d_future, fbytes, cudaMemcpyDeviceToHost, m->stream));
        if (direct) m->sparse_rows.invalidate();  // a dense copy went into the buffer: the row list no longer describes it
    }
    if (sparse) m->sparse_rows.clear_previous(future, mc.V, mc.T);  // host work that needs no new data: done while the device finishes the frame
    CK(cudaStreamSynchronize(m->stream));
    if (sparse) {
        const int nf = *m->h_nf;
        if (nf > fguess) {  // more rows than the speculative copy carried
            CK(cudaMemcpyAsync(m->h_fidx + fguess, m->d_fidx + fguess, sizeof(int) * (size_t)(nf - fguess), cudaMemcpyDeviceToHost, m->stream));
            CK(cudaMemcpyAsync(m->h_fval + (size_t)fguess * mc.T, m->d_fval + (size_t)fguess * mc.T, sizeof(float) * (size_t)(nf - fguess) * mc.T,
                               cudaMemcpyDeviceToHost, m->stream));
            CK(cudaStreamSynchronize(m->stream));
        }
        m->fut_guess = nf + nf / 4 + 1024;  // the row count moves by a few per cent from frame to frame; twice the count doubled the copy
        m->sparse_rows.apply(future, mc.V, mc.T, m->h_fidx, m->h_fval, nf);
        m->last_d2h_bytes = 4 + (long long)std::max(nf, fguess) * (4 + 4 * mc.T);
    } else {
        m->last_d2h_bytes = future ? (long long)fbytes : 0;
    }
    if (m->update_counter > 0 && m->state_event_recorded) absorb_state(m);  // the frame's state copy has landed by now
    int n = sparse ? m->h_nf[1] : *m->h_count;
    if (n_out) *n_out = n;
    int ncopy = std::min(n, cap);
    m->last_d2h_bytes += 4 + 12ll * std::max(ncopy, guess);
    m->occ_guess = n + n / 4 + 512;
    if (xyz_out && ncopy > 0) {
        if (ncopy > guess) {
            CK(cudaMemcpyAsync(m->h_xyz + 3 * (size_t)guess, m->d_xyz + 3 * (size_t)guess, sizeof(float) * 3 * (size_t)(ncopy - guess),
                               cudaMemcpyDeviceToHost, m->stream));
            CK(cudaStreamSynchronize(m->stream));
        }
        memcpy(xyz_out, m->h_xyz, sizeof(float) * 3 * (size_t)ncopy);
    }
    if (future && !direct) memcpy(future, m->h_future, fbytes);
    if (m->profile) prof_collect(m);
    return DSPMAP_OK;
}
// Pipelined reader: the reader kernels run on the map's stream, the device-to-host copies on a second stream, so they
// overlap the next update.  Two result slots alternate; a slot's host pointers stay valid until it is reused.
int dspmap_get_occupancy_async(dspmap *m, float thr, int with_future, int *ticket) {
    if (!m || !ticket) return DSPMAP_E_BAD_ARG;
    CK(cudaSetDevice(m->cfg.device));
    const MapConst &mc = m->mc;
    const size_t fcount = (size_t)mc.V * std::max(mc.T, 1);
    if (!m->copy_stream) {
        CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&m->ev_reader, cudaEventDisableTiming));
        for (auto &s : m->rslot) {
            if (dalloc(m, &s.d_xyz, (size_t)mc.V * 3, false) != DSPMAP_OK || dalloc(m, &s.d_future, fcount, false) != DSPMAP_OK ||
                dalloc(m, &s.d_count, 1) != DSPMAP_OK)
                return DSPMAP_E_CUDA;
            CK(cudaMallocHost(&s.h_xyz, sizeof(float) * 3 * (size_t)mc.V));
            CK(cudaMallocHost(&s.h_future, sizeof(float) * fcount));
            CK(cudaMallocHost(&s.h_count, sizeof(int)));
            CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
    }
    const int k = m->next_slot;
    m->next_slot ^= 1;
    auto &s = m->rslot[k];
    if (s.pending) CK(cudaStreamWaitEvent(m->stream, s.done, 0));  // the slot's previous copies must have left the device buffers
    int rc = dspmap_get_occupancy_device(m, thr, s.d_xyz, mc.V, s.d_count, with_future ? s.d_future : nullptr);
    if (rc != DSPMAP_OK) return rc;
    CK(cudaEventRecord(m->ev_reader, m->stream));
    CK(cudaStreamWaitEvent(m->copy_stream, m->ev_reader, 0));
    s.copied = std::min(m->occ_guess, mc.V);
    s.with_future = with_future != 0;
    CK(cudaMemcpyAsync(s.h_count, s.d_count, sizeof(int), cudaMemcpyDeviceToHost, m->copy_stream));
    CK(cudaMemcpyAsync(s.h_xyz, s.d_xyz, sizeof(float) * 3 * (size_t)s.copied, cudaMemcpyDeviceToHost, m->copy_stream));
    if (with_future) CK(cudaMemcpyAsync(s.h_future, s.d_future, sizeof(float) * (size_t)mc.V * mc.T, cudaMemcpyDeviceToHost, m->copy_stream));
    CK(cudaEventRecord(s.done, m->copy_stream));
    s.pending = true;
    *ticket = k;
    return DSPMAP_OK;
}
int dspmap_wait_occupancy(dspmap *m, int ticket, const float **xyz, int *n_out, const float **future) {
    if (!m || ticket < 0 || ticket > 1 || !m->rslot[ticket].pending) { g_err = "no such pending reader ticket"; return DSPMAP_E_BAD_ARG; }
    CK(cudaSetDevice(m->cfg.device));
    auto &s = m->rslot[ticket];
    CK(cudaEventSynchronize(s.done));
    const int n = *s.h_count;
    if (n > s.copied) {  // more occupied voxels than the speculative copy carried
        CK(cudaMemcpyAsync(s.h_xyz + 3 * (size_t)s.copied, s.d_xyz + 3 * (size_t)s.copied, sizeof(float) * 3 * (size_t)(n - s.copied),
                           cudaMemcpyDeviceToHost, m->copy_stream));
        CK(cudaStreamSynchronize(m->copy_stream));
        s.copied = n;
    }
    m->occ_guess = n + n / 4 + 512;
    if (n_out) *n_out = n;
    if (xyz) *xyz = s.h_xyz;
    if (future) *future = s.with_future ? s.h_future : nullptr;
    if (m->profile) prof_collect(m);
    return DSPMAP_OK;
}
long long dspmap_last_reader_bytes(dspmap *m) { return m ? m->last_d2h_bytes : 0; }
int dspmap_timeline(dspmap *m, float *out8) {
    if (!m || !out8 || !m->timeline) return DSPMAP_E_BAD_ARG;
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    cudaEvent_t evs[7] = {m->ev_feat, m->ev_arrived, m->ev_join_obs, m->ev_nb_early, m->ev_fork, m->ev_join, m->ev_state};
    for (int k = 0; k < 7; ++k) {
        out8[k] = -1.f;
        if (evs[k] && cudaEventElapsedTime(&out8[k], m->ev_fork_obs, evs[k]) != cudaSuccess) { out8[k] = -1.f; cudaGetLastError(); }
    }
    out8[7] = 0.f;
    return DSPMAP_OK;
}
int dspmap_estimator_stats(dspmap *m, int32_t *out8) {
    if (!m || !out8) return DSPMAP_E_BAD_ARG;
    for (int c = 0; c < 7; ++c) out8[c] = m->est_h_hdr ? m->est_h_hdr[c] : 0;
    out8[7] = m->est_h_hdr ? m->est_h_hdr[EC_NTAGGED] : 0;
    return m->est_gpu ? 1 : 0;
}
void dspmap_last_update_bytes(dspmap *m, long long *h2d, long long *d2h) {
    if (h2d) *h2d = m ? m->upd_h2d_bytes : 0;
    if (d2h) *d2h = m ? m->upd_d2h_bytes : 0;
}
int dspmap_pin_host_buffer(dspmap *m, void *ptr, size_t bytes) {
    if (!m) return DSPMAP_E_BAD_ARG;
    if (m->pinned_user) {
        cudaHostUnregister(m->pinned_user);
        m->pinned_user = nullptr;
        m->pinned_bytes = 0;
    }
    m->sparse_rows.invalidate();
    if (ptr && bytes) {
        if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) {
            cudaGetLastError();  // not fatal: the staged path stays in use
            return DSPMAP_REJECTED;
        }
        m->pinned_user = ptr;
        m->pinned_bytes = bytes;
    }
    return DSPMAP_OK;
}
int dspmap_clear_prediction(dspmap *m) {
    if (!m) return DSPMAP_E_BAD_ARG;
    LAUNCH(m, FAM_READER, k_future_clear, kSMs * 4, 256, 0, m->mc, m->dp);
    CK(cudaStreamSynchronize(m->stream));
    return DSPMAP_OK;
}
int dspmap_get_tagged_cloud(dspmap *m, float *out, int cap) {
    if (!m) return DSPMAP_E_BAD_ARG;
    int rc = fetch_tagged(m);
    if (rc != DSPMAP_OK) return rc;
    int n = (int)(m->tagged_host.size() / 7);
    if (out) memcpy(out, m->tagged_host.data(), sizeof(float) * 7 * (size_t)std::min(n, cap));
    return n;
}

void dspmap_voxel_center(const dspmap *m, int index, float *xyz) { dsp_voxel_center(m->mc, index, xyz); }
int dspmap_voxel_index(const dspmap *m, float x, float y, float z, int *index) {
    int i = dsp_voxel_index(m->mc, x, y, z);
    if (i < 0) return 0;
    *index = i;
    return 1;
}
float dspmap_uniform(dspmap *m, float lo, float hi) { return m->estimator.uniform(lo, hi); }

void dspmap_dims(const dspmap *m, int32_t *d) {
    const MapConst &mc = m->mc;
    int v[] = {mc.V, mc.S, mc.P, mc.L, mc.T, mc.Nh, mc.Nv, mc.NBW, mc.max_ppv, mc.nx, mc.ny, mc.nz, mc.OBS, mc.model, mc.fast_res, m->fast_sigma};
    memcpy(d, v, sizeof(v));
}

int dspmap_dump_particles(dspmap *m, int32_t *ids, float *vals, int cap) {
    if (!m) return DSPMAP_E_BAD_ARG;
    CK(cudaSetDevice(m->cfg.device));
    const MapConst &mc = m->mc;
    LAUNCH(m, FAM_MISC, k_reset_counter, 1, 1, 0, &m->dp.st->n_live);
    LAUNCH(m, FAM_MISC, k_enumerate, grid_for(mc.V, 256), 256, 0, mc, m->dp, 0);
    int n = 0;
    CK(cudaMemcpyAsync(&n, &m->dp.st->n_live, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (!ids || n == 0) return n;
    int *d_keys;
    float *d_vals;
    CK(cudaMalloc(&d_keys, sizeof(int) * (size_t)n));
    CK(cudaMalloc(&d_vals, sizeof(float) * 8 * (size_t)n));
    LAUNCH(m, FAM_MISC, k_dump_gather, grid_for(n, 256), 256, 0, mc, m->dp, d_keys, d_vals);
    std::vector<int> keys(n);
    std::vector<float> v(8 * (size_t)n);
    CK(cudaMemcpyAsync(keys.data(), d_keys, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaMemcpyAsync(v.data(), d_vals, sizeof(float) * 8 * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    cudaFree(d_keys);
    cudaFree(d_vals);
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
    for (int i = 0; i < n && i < cap; ++i) {
        int k = keys[order[i]];
        ids[2 * i] = k >> DSP_KEY_SHIFT;
        ids[2 * i + 1] = k & (DSP_MAX_SLOTS - 1);
        memcpy(vals + 8 * (size_t)i, &v[8 * (size_t)order[i]], 8 * sizeof(float));
    }
    return n;
}
int dspmap_load_particles(dspmap *m, const int32_t *ids, const float *vals, int n) {
    if (!m || n < 0) return DSPMAP_E_BAD_ARG;
    CK(cudaSetDevice(m->cfg.device));
    m->vz_mode = false;
    for (int i = 0; i < n; ++i) {
        const float *r = vals + 8 * (size_t)i;
        if (r[3] != 0.f || !(r[0] > 0.1f && r[0] < 6.f)) { m->vz_mode = true; break; }
    }
    return upload_particles(m, ids, vals, n);
}
int dspmap_dump_voxel_objects(dspmap *m, float *out) {
    if (!m) return DSPMAP_E_BAD_ARG;
    const MapConst &mc = m->mc;
    std::vector<float> occ(4 * (size_t)mc.V), fut((size_t)mc.V * std::max(mc.T, 1));
    CK(cudaStreamSynchronize(m->stream));  // the map's stream does not synchronise with the default stream
    CK(cudaMemcpy(occ.data(), m->dp.OCCV, sizeof(float) * occ.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fut.data(), m->dp.FUT, sizeof(float) * fut.size(), cudaMemcpyDeviceToHost));
    for (int v = 0; v < mc.V; ++v) {
        float *o = out + (size_t)v * (4 + mc.T);
        memcpy(o, &occ[4 * (size_t)v], 16);
        for (int t = 0; t < mc.T; ++t) o[4 + t] = fut[(size_t)v * mc.T + t];
    }
    return DSPMAP_OK;
}
int dspmap_dump_observations(dspmap *m, int32_t *counts, float *maxlen, float *pts) {
    if (!m) return DSPMAP_E_BAD_ARG;
    const MapConst &mc = m->mc;
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaMemcpy(counts, m->dp.obs_cnt, sizeof(int) * mc.P, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(maxlen, m->dp.obs_maxbits, sizeof(int) * mc.P, cudaMemcpyDeviceToHost));
    for (int i = 0; i < mc.P; ++i) counts[i] = std::min(counts[i], mc.OBS - 1);  // :282-284
    if (pts) {
        std::vector<float> o(4 * (size_t)mc.P * mc.OBS), cz((size_t)mc.P * mc.OBS);
        CK(cudaMemcpy(o.data(), m->dp.OBSP, sizeof(float) * o.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cz.data(), m->dp.CZ, sizeof(float) * cz.size(), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < (size_t)mc.P * mc.OBS; ++k) {
            pts[5 * k] = o[4 * k]; pts[5 * k + 1] = o[4 * k + 1]; pts[5 * k + 2] = o[4 * k + 2];
            pts[5 * k + 3] = cz[k];
            pts[5 * k + 4] = o[4 * k + 3];
        }
    }
    return DSPMAP_OK;
}
int dspmap_dump_pyramid_lists(dspmap *m, int32_t *offsets, int32_t *entries, int cap) {
    if (!m) return DSPMAP_E_BAD_ARG;
    const MapConst &mc = m->mc;
    CK(cudaStreamSynchronize(m->stream));
    std::vector<int> poff(mc.P + 1), plen(mc.P);
    CK(cudaMemcpy(poff.data(), m->dp.poff, sizeof(int) * (mc.P + 1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(plen.data(), m->dp.plen, sizeof(int) * mc.P, cudaMemcpyDeviceToHost));
    std::vector<int> la(std::max(poff[mc.P], 1));
    CK(cudaMemcpy(la.data(), m->dp.LA, sizeof(int) * (size_t)poff[mc.P], cudaMemcpyDeviceToHost));
    int n = 0;
    for (int p = 0; p < mc.P; ++p) {
        offsets[p] = n;
        for (int j = 0; j < plen[p]; ++j) {
            if (entries && n < cap) {
                int a = la[poff[p] + j];
                entries[2 * n] = a / mc.S;
                entries[2 * n + 1] = a % mc.S;
            }
            ++n;
        }
    }
    offsets[mc.P] = n;
    return n;
}
int dspmap_dump_plane_normals(dspmap *m, float *h, float *v) {
    if (!m || !h || !v) return DSPMAP_E_BAD_ARG;
    const MapConst &mc = m->mc;
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaMemcpy(h, m->dp.planes, sizeof(float) * 3 * (mc.Nh + 1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v, m->dp.planes + 3 * (mc.Nh + 1), sizeof(float) * 3 * (mc.Nv + 1), cudaMemcpyDeviceToHost));
    return DSPMAP_OK;
}
int dspmap_cursors(dspmap *m, int64_t *c) {
    if (!m) return DSPMAP_E_BAD_ARG;
    DevState st;
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaMemcpy(&st, m->dp.st, sizeof(st), cudaMemcpyDeviceToHost));
    c[0] = st.p_cur; c[1] = st.v_cur; c[2] = st.u_cur;
    return DSPMAP_OK;
}
int dspmap_set_cursors(dspmap *m, int64_t p, int64_t v, int64_t u) {
    if (!m) return DSPMAP_E_BAD_ARG;
    DevState st;
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaMemcpy(&st, m->dp.st, sizeof(st), cudaMemcpyDeviceToHost));
    st.p_cur = p; st.v_cur = v; st.u_cur = u;
    CK(cudaMemcpy(m->dp.st, &st, sizeof(st), cudaMemcpyHostToDevice));
    return DSPMAP_OK;
}
int dspmap_counters(dspmap *m, int64_t *out) {
    if (!m) return DSPMAP_E_BAD_ARG;
    if (m->update_counter > 0 && m->state_event_recorded) {  // the last frame may still be running: wait for its state copy
        CK(cudaStreamSynchronize(m->stream));
        absorb_state(m);
    }
    const DevState &s = m->last_state;
    int64_t v[16] = {s.n_live, s.n_left_map, s.n_voxel_full, s.n_pyramid_full, s.n_moved, s.n_fov, s.n_cand, s.n_born,
                     s.n_low_weight, s.n_pre, s.n_old, s.n_out, s.n_valid, (int64_t)(s.overflow | m->overflow_latched), m->launches_frame, m->launches_total};
    memcpy(out, v, sizeof(v));
    return DSPMAP_OK;
}
int dspmap_set_last_pose(dspmap *m, float px, float py, float pz, double t) {
    if (!m) return DSPMAP_E_BAD_ARG;
    m->last_p[0] = px; m->last_p[1] = py; m->last_p[2] = pz;
    m->last_t = t;
    m->have_last = true;
    return DSPMAP_OK;
}
int dspmap_set_stage_limit(dspmap *m, int k) { if (!m) return DSPMAP_E_BAD_ARG; m->stage_limit = k; return DSPMAP_OK; }
int dspmap_set_stream(dspmap *m, void *s) {
    if (!m) return DSPMAP_E_BAD_ARG;
    cudaStreamSynchronize(m->stream);
    m->stream = s ? (cudaStream_t)s : m->own_stream;
    return DSPMAP_OK;
}
int dspmap_synchronize(dspmap *m) {
    if (!m) return DSPMAP_E_BAD_ARG;
    CK(cudaStreamSynchronize(m->stream));
    if (m->profile) prof_collect(m);
    if (m->update_counter > 0 && m->state_event_recorded) absorb_state(m);  // every frame ends with an async copy of the device state
    return report_overflow(m);
}
int dspmap_profile_enable(dspmap *m, int on) {
    if (!m) return DSPMAP_E_BAD_ARG;
    cudaStreamSynchronize(m->stream);
    prof_collect(m);
    m->profile = on != 0;
    for (int i = 0; i < FAM_COUNT; ++i) { m->prof_ms[i] = 0; m->prof_n[i] = 0; }
    m->prof_kernel.clear();
    return DSPMAP_OK;
}
int dspmap_profile_read(dspmap *m, const char **names, float *ms, int32_t *launches, int cap) {
    if (!m) return DSPMAP_E_BAD_ARG;
    cudaStreamSynchronize(m->stream);
    prof_collect(m);
    int n = std::min<int>(cap, FAM_COUNT);
    for (int i = 0; i < n; ++i) {
        names[i] = kFamilyNames[i];
        ms[i] = (float)m->prof_ms[i];
        launches[i] = m->prof_n[i];
    }
    return n;
}

// Same measurement per kernel name (the name as written at the launch site): names[i] stays valid until the next call.
int dspmap_profile_read_kernels(dspmap *m, const char **names, float *ms, int32_t *launches, int cap) {
    if (!m) return DSPMAP_E_BAD_ARG;
    cudaStreamSynchronize(m->stream);
    prof_collect(m);
    int n = 0;
    for (auto &kv : m->prof_kernel) {
        if (n >= cap) break;
        names[n] = kv.first.c_str();
        ms[n] = (float)kv.second.first;
        launches[n] = kv.second.second;
        ++n;
    }
    return n;
}

struct dspmap_estimator {
    MapConst mc;
    int model;
    std::vector<float> planes0, tagged;
    VelocityEstimator est;
    bool threaded = false;
    HostWorker worker;
};
dspmap_estimator *dspmap_estimator_create(const dspmap_config *cfg, float filter_res) {
    dspmap_estimator *e = new dspmap_estimator();
    memset(&e->mc, 0, sizeof(e->mc));
    e->mc.Nh = cfg->half_fov_h * 2 / cfg->angle_resolution;
    e->mc.Nv = cfg->half_fov_v * 2 / cfg->angle_resolution;
    e->model = cfg->model;
    make_planes0(cfg, e->mc.Nh, e->mc.Nv, e->planes0);
    e->est.reset(cfg->uniform_seed);
    e->est.filter_res = filter_res;
    return e;
}
void dspmap_estimator_destroy(dspmap_estimator *e) { delete e; }
int dspmap_estimator_set_threaded(dspmap_estimator *e, int on) {
    if (!e) return DSPMAP_E_BAD_ARG;
    e->threaded = on != 0;
    if (e->threaded) e->worker.start(); else e->worker.stop();
    return DSPMAP_OK;
}
int dspmap_estimator_estimate(dspmap_estimator *e, int n, const float *pts, float px, float py, float pz, float dt,
                              float qw, float qx, float qy, float qz, float *out, int cap) {
    FrameConst fc;
    memset(&fc, 0, sizeof(fc));
    fc.q[0] = qw; fc.q[1] = qx; fc.q[2] = qy; fc.q[3] = qz;
    dsp_quat_inverse(fc.q, fc.qi);
    fc.cur[0] = px; fc.cur[1] = py; fc.cur[2] = pz;
    fc.dt = dt;
    size_t before = e->tagged.size();
    std::vector<float> prev;
    prev.swap(e->tagged);
    e->tagged.assign(1, -12345.f);  // sentinel: estimate() leaves the vector untouched when nothing is in view
    if (e->threaded) {  // the same hand-over dspmap_update uses with DSPMAP_EST_THREAD=1
        e->worker.submit([e, fc, pts, n] { e->est.estimate(e->mc, fc, e->planes0.data(), pts, n, e->model, e->tagged); });
        e->worker.wait();
    } else {
        e->est.estimate(e->mc, fc, e->planes0.data(), pts, n, e->model, e->tagged);
    }
    if (e->tagged.size() == 1 && e->tagged[0] == -12345.f) {
        e->tagged.swap(prev);
        (void)before;
        return -1;
    }
    int nt = (int)(e->tagged.size() / 7);
    if (out) memcpy(out, e->tagged.data(), sizeof(float) * 7 * (size_t)std::min(nt, cap));
    return nt;
}

int dspmap_euclidean_clusters(const float *xyz, int n, float tolerance, int min_size, int max_size, int path, int *labels) {
    if (n < 0 || (n > 0 && (!xyz || !labels))) return DSPMAP_E_BAD_ARG;
    std::vector<std::vector<int>> cl;
    if (!euclidean_clusters_path(xyz, n, tolerance, min_size, max_size, path, cl)) return DSPMAP_E_CAPACITY;
    for (int i = 0; i < n; ++i) labels[i] = -1;
    for (size_t c = 0; c < cl.size(); ++c)
        for (int idx : cl[c]) labels[idx] = (int)c;
    return (int)cl.size();
}

}  // extern "C"

#include "shard_host.inc"

namespace {
// the reference's particle CSV (dsp_dynamic.h:328-350): flag,vx,vy,vz,px,py,pz,weight,voxel per live particle
int write_particle_csv(dspmap *m) {
    int n = dspmap_dump_particles(m, nullptr, nullptr, 0);
    if (n < 0) return n;
    std::vector<int32_t> ids(2 * (size_t)std::max(n, 1));
    std::vector<float> vals(8 * (size_t)std::max(n, 1));
    n = dspmap_dump_particles(m, ids.data(), vals.data(), n);
    std::string name = m->record_prefix + "particles_update_t_" + std::to_string(m->update_counter) + "_" +
                       std::to_string((int)(m->update_time * 1000)) + ".csv";
    std::ofstream f(name, std::ios::out | std::ios::trunc);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 8; ++k) f << vals[8 * (size_t)i + k] << ",";
        f << ids[2 * i] << "\n";
    }
    return DSPMAP_OK;
}
}  // namespace
