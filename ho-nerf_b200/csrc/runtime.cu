// Library runtime: error text, launch counter, device info, weight-norm pack/backward, column sums.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "fields_common.cuh"

namespace hn {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_timing_on{0};
static std::mutex g_timing_mu;
static std::vector<cudaEvent_t> g_ev_start, g_ev_stop;
static std::vector<int> g_ev_tag;
static size_t g_ev_used = 0;

TimingScope::TimingScope(cudaStream_t s, int tag) : stream(s), slot(-1) {
    if (!g_timing_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (g_ev_used == g_ev_start.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        g_ev_start.push_back(a);
        g_ev_stop.push_back(b);
        g_ev_tag.push_back(0);
    }
    slot = (int)g_ev_used++;
    g_ev_tag[slot] = tag;
    cudaEventRecord(g_ev_start[slot], stream);
}
TimingScope::~TimingScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    cudaEventRecord(g_ev_stop[slot], stream);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

// ---- weight norm ---------------------------------------------------------------------------
// one warp per output row
// packed column of input column k when `gap` zero columns follow input column gap_at
__device__ __forceinline__ int packed_col(int k, int gap_at, int gap) { return k < gap_at ? k : k + gap; }

__global__ void wn_pack_kernel(const float* __restrict__ v, const float* __restrict__ g, int out_dim,
                               int in_dim, int ld, float post_scale, float* __restrict__ W,
                               float* __restrict__ WT, int ldT, int gap_at, int gap) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= ldT && row >= out_dim) return;
    if (row >= out_dim) {            // zero padding columns of the transposed copy
        if (WT && row < ldT)
            for (int k = lane; k < in_dim + gap; k += 32) WT[(int64_t)k * ldT + row] = 0.0f;
        return;
    }
    const float* vr = v + (int64_t)row * in_dim;
    float ss = 0.0f;
    for (int k = lane; k < in_dim; k += 32) ss = fmaf(vr[k], vr[k], ss);
    ss = warp_sum(ss);
    float sc = post_scale * (g[row] / sqrtf(ss));
    float* wr = W + (int64_t)row * ld;
    for (int k = lane; k < ld; k += 32) {         // k = packed column
        int src = k < gap_at ? k : k - gap;
        bool live = (k < gap_at || k >= gap_at + gap) && src < in_dim;
        float w = live ? vr[src] * sc : 0.0f;
        wr[k] = w;
        if (WT && k < in_dim + gap) WT[(int64_t)k * ldT + row] = w;
    }
}

__global__ void wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                              const float* __restrict__ dW, int out_dim, int in_dim, int ld,
                              float post_scale, float* __restrict__ dv, float* __restrict__ dg,
                              int gap_at, int gap) {
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= out_dim) return;
    const float* vr = v + (int64_t)row * in_dim;
    const float* dr = dW + (int64_t)row * ld;
    float ss = 0.0f, dot = 0.0f;
    for (int k = lane; k < in_dim; k += 32) {
        ss = fmaf(vr[k], vr[k], ss);
        dot = fmaf(dr[packed_col(k, gap_at, gap)], vr[k], dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot) * post_scale;
    float nrm = sqrtf(ss);
    float gn = g[row] / nrm;
    if (lane == 0) dg[row] = dot / nrm;
    float coef = dot / ss;
    float* dvr = dv + (int64_t)row * in_dim;
    for (int k = lane; k < in_dim; k += 32) dvr[k] = gn * (post_scale * dr[packed_col(k, gap_at, gap)] - coef * vr[k]);
}

// ---- all layers of a net in one launch --------------------------------------------------------------------
struct WnJobs {
    int n;
    hn_wn_job_t job[HN_MAX_LAYERS];
};
// one warp per output row; blockIdx.y = layer
__global__ void wn_pack_batch_kernel(const __grid_constant__ WnJobs jobs) {
    const hn_wn_job_t& j = jobs.job[blockIdx.y];
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= j.ldT && row >= j.out_dim) return;
    if (row >= j.out_dim) {
        if (j.WT && row < j.ldT)
            for (int k = lane; k < j.in_dim + j.gap; k += 32) j.WT[(int64_t)k * j.ldT + row] = 0.0f;
        return;
    }
    const float* vr = j.v + (int64_t)row * j.in_dim;
    float ss = 0.0f;
    for (int k = lane; k < j.in_dim; k += 32) ss = fmaf(vr[k], vr[k], ss);
    ss = warp_sum(ss);
    const float sc = j.post_scale * (j.g[row] / sqrtf(ss));
    float* wr = j.W + (int64_t)row * j.ld;
    for (int k = lane; k < j.ld; k += 32) {
        const int src = k < j.gap_at ? k : k - j.gap;
        const bool live = (k < j.gap_at || k >= j.gap_at + j.gap) && src < j.in_dim;
        const float w = live ? vr[src] * sc : 0.0f;
        wr[k] = w;
        if (j.WT && k < j.in_dim + j.gap) j.WT[(int64_t)k * j.ldT + row] = w;
    }
}
__global__ void wn_bwd_batch_kernel(const __grid_constant__ WnJobs jobs) {
    const hn_wn_job_t& j = jobs.job[blockIdx.y];
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= j.out_dim) return;
    const float* vr = j.v + (int64_t)row * j.in_dim;
    const float* dr = j.dW + (int64_t)row * j.ld;
    float ss = 0.0f, dot = 0.0f;
    for (int k = lane; k < j.in_dim; k += 32) {
        ss = fmaf(vr[k], vr[k], ss);
        dot = fmaf(dr[packed_col(k, j.gap_at, j.gap)], vr[k], dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot) * j.post_scale;
    const float nrm = sqrtf(ss);
    const float gn = j.g[row] / nrm;
    if (lane == 0) j.dg[row] = dot / nrm;
    const float coef = dot / ss;
    float* dvr = j.dv + (int64_t)row * j.in_dim;
    for (int k = lane; k < j.in_dim; k += 32)
        dvr[k] = gn * (j.post_scale * dr[packed_col(k, j.gap_at, j.gap)] - coef * vr[k]);
}

// ---- column sums -----------------------------------------------------------------------------
// block = 32 columns x 8 row lanes; grid.y splits the rows
__global__ void colsum_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, int cols,
                              float scale, int64_t rows_per_block, float* __restrict__ out) {
    __shared__ float red[8][33];
    int c = blockIdx.x * 32 + (threadIdx.x & 31);
    int rl = threadIdx.x >> 5;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = min(rows, r0 + rows_per_block);
    float acc = 0.0f;
    if (c < cols)
        for (int64_t r = r0 + rl; r < r1; r += 8) acc += X[r * ldx + c];
    red[rl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(&out[c], t * scale);
    }
}

int launch_colsum(const float* X, int64_t ldx, int64_t rows, int cols, float scale, float* out,
                  cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return HN_OK;
    int64_t rpb = 1024;
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, rpb));
    colsum_kernel<<<grid, 256, 0, stream>>>(X, ldx, rows, cols, scale, rpb, out);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace hn

using namespace hn;

extern "C" {

const char* hn_last_error(void) { return g_err; }
int hn_version(void) { return 100; }
int64_t hn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int hn_timing_enable(int on) { g_timing_on.store(on ? 1 : 0); return HN_OK; }
int hn_timing_reset(void) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_ev_used = 0;
    return HN_OK;
}
int hn_timing_collect(double* total_ms, int64_t* n_launches) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    double tot = 0.0;
    for (size_t i = 0; i < g_ev_used; ++i) {
        float ms = 0.0f;
        HN_CHECK_CUDA(cudaEventSynchronize(g_ev_stop[i]));
        HN_CHECK_CUDA(cudaEventElapsedTime(&ms, g_ev_start[i], g_ev_stop[i]));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (n_launches) *n_launches = (int64_t)g_ev_used;
    return HN_OK;
}

int hn_timing_collect_tags(double* ms_per_tag, int64_t* launches_per_tag, int n_tags) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (int t = 0; t < n_tags; ++t) {
        if (ms_per_tag) ms_per_tag[t] = 0.0;
        if (launches_per_tag) launches_per_tag[t] = 0;
    }
    for (size_t i = 0; i < g_ev_used; ++i) {
        float ms = 0.0f;
        HN_CHECK_CUDA(cudaEventSynchronize(g_ev_stop[i]));
        HN_CHECK_CUDA(cudaEventElapsedTime(&ms, g_ev_start[i], g_ev_stop[i]));
        const int t = g_ev_tag[i];
        if (t >= 0 && t < n_tags) {
            if (ms_per_tag) ms_per_tag[t] += ms;
            if (launches_per_tag) launches_per_tag[t] += 1;
        }
    }
    return HN_OK;
}

int hn_wn_pack_batch(const hn_wn_job_t* jobs, int n, hn_stream_t stream) {
    HN_REQUIRE(jobs && n >= 1 && n <= HN_MAX_LAYERS, "hn_wn_pack_batch: bad job list");
    WnJobs w;
    w.n = n;
    int max_rows = 0;
    for (int i = 0; i < n; ++i) {
        const hn_wn_job_t& j = jobs[i];
        HN_REQUIRE(j.v && j.g && j.W && j.out_dim > 0 && j.in_dim > 0 && j.gap >= 0 && j.gap_at >= 0 && j.gap_at <= j.in_dim &&
                       j.ld >= j.in_dim + j.gap, "hn_wn_pack_batch: bad arguments for layer %d", i);
        HN_REQUIRE(!j.WT || j.ldT >= j.out_dim, "hn_wn_pack_batch: ldT too small for layer %d", i);
        w.job[i] = j;
        max_rows = std::max(max_rows, std::max(j.out_dim, j.WT ? j.ldT : 0));
    }
    wn_pack_batch_kernel<<<dim3((unsigned)ceil_div((int64_t)max_rows * 32, 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(w);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_wn_bwd_batch(const hn_wn_job_t* jobs, int n, hn_stream_t stream) {
    HN_REQUIRE(jobs && n >= 1 && n <= HN_MAX_LAYERS, "hn_wn_bwd_batch: bad job list");
    WnJobs w;
    w.n = n;
    int max_rows = 0;
    for (int i = 0; i < n; ++i) {
        const hn_wn_job_t& j = jobs[i];
        HN_REQUIRE(j.v && j.g && j.dW && j.dv && j.dg && j.out_dim > 0 && j.in_dim > 0 && j.ld >= j.in_dim + j.gap,
                   "hn_wn_bwd_batch: bad arguments for layer %d", i);
        w.job[i] = j;
        max_rows = std::max(max_rows, j.out_dim);
    }
    wn_bwd_batch_kernel<<<dim3((unsigned)ceil_div((int64_t)max_rows * 32, 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(w);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_wn_pack(const float* v, const float* g, int out_dim, int in_dim, int ld, float post_scale,
               float* W, float* WT, int ldT, hn_stream_t stream) {
    return hn_wn_pack_gap(v, g, out_dim, in_dim, ld, post_scale, in_dim, 0, W, WT, ldT, stream);
}

int hn_wn_pack_gap(const float* v, const float* g, int out_dim, int in_dim, int ld, float post_scale,
                   int gap_at, int gap, float* W, float* WT, int ldT, hn_stream_t stream) {
    HN_REQUIRE(v && g && W && out_dim > 0 && in_dim > 0 && gap >= 0 && gap_at >= 0 && gap_at <= in_dim &&
                   ld >= in_dim + gap, "hn_wn_pack: bad arguments");
    HN_REQUIRE(!WT || ldT >= out_dim, "hn_wn_pack: ldT too small");
    int rows = WT ? (ldT > out_dim ? ldT : out_dim) : out_dim;
    wn_pack_kernel<<<(unsigned)ceil_div((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        v, g, out_dim, in_dim, ld, post_scale, W, WT, WT ? ldT : 0, gap_at, gap);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_wn_bwd(const float* v, const float* g, const float* dW, int out_dim, int in_dim, int ld,
              float post_scale, float* dv, float* dg, hn_stream_t stream) {
    return hn_wn_bwd_gap(v, g, dW, out_dim, in_dim, ld, post_scale, in_dim, 0, dv, dg, stream);
}

int hn_wn_bwd_gap(const float* v, const float* g, const float* dW, int out_dim, int in_dim, int ld,
                  float post_scale, int gap_at, int gap, float* dv, float* dg, hn_stream_t stream) {
    HN_REQUIRE(v && g && dW && dv && dg && out_dim > 0 && in_dim > 0 && gap >= 0 && ld >= in_dim + gap,
               "hn_wn_bwd: bad arguments");
    wn_bwd_kernel<<<(unsigned)ceil_div((int64_t)out_dim * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        v, g, dW, out_dim, in_dim, ld, post_scale, dv, dg, gap_at, gap);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
