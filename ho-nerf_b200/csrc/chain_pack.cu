// Weight packing for the fused chain kernels: fp32 effective weights -> bf16 hi/lo tiles already in the
// tcgen05 shared-memory layout (K-major, SWIZZLE_128B), so the kernels' producer warp streams them with
// plain bulk copies.  Runs once per parameter version.
#include <algorithm>

#include "chain_common.cuh"

namespace hn {
namespace chain {

__global__ void pack_b_kernel(const float* __restrict__ src, int64_t ld, PackMap m, int rows, int cols, int n_pad,
                              int kblocks, uint8_t* __restrict__ dst, int lo16) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;          // one 16-byte chunk (8 columns) of one row
    int total = n_pad * kblocks * 8;
    if (idx >= total) return;
    int n = idx / (kblocks * 8);
    int c = idx - n * (kblocks * 8);
    int kb = c >> 3, c16 = c & 7;
    const int64_t srow = n < m.rsplit ? m.row0 + n : m.row1 + (n - m.rsplit);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int k = kb * 64 + c16 * 8 + j;
        const int scol = k < m.ksplit ? m.col0 + k : m.col1 + (k - m.ksplit);
        v[j] = (n < rows && k < cols) ? src[srow * ld + scol] : 0.0f;
    }
    uint4 hi, lo;
    if (lo16) {
        split2_lo16(v[0], v[1], hi.x, lo.x);
        split2_lo16(v[2], v[3], hi.y, lo.y);
        split2_lo16(v[4], v[5], hi.z, lo.z);
        split2_lo16(v[6], v[7], hi.w, lo.w);
    } else {
        split2(v[0], v[1], hi.x, lo.x);
        split2(v[2], v[3], hi.y, lo.y);
        split2(v[4], v[5], hi.z, lo.z);
        split2(v[6], v[7], hi.w, lo.w);
    }
    size_t base = (size_t)kb * 2 * n_pad * 128 + tc::sw128_offset((uint32_t)n, (uint32_t)c16);
    *reinterpret_cast<uint4*>(dst + base) = hi;
    *reinterpret_cast<uint4*>(dst + base + (size_t)n_pad * 128) = lo;
}

struct PackJob {
    const float* src;
    int64_t ld;
    PackMap m;
    int rows, cols, n_pad, kblocks;
    uint8_t* dst;
    int lo16;
};
struct PackJobs {
    int n;
    PackJob job[PACK_MAX_JOBS];
};

__global__ void pack_b_batch_kernel(const __grid_constant__ PackJobs jobs) {
    const PackJob& j = jobs.job[blockIdx.y];
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int total = j.n_pad * j.kblocks * 8;
    if (idx >= total) return;
    int n = idx / (j.kblocks * 8);
    int c = idx - n * (j.kblocks * 8);
    int kb = c >> 3, c16 = c & 7;
    const int64_t srow = n < j.m.rsplit ? j.m.row0 + n : j.m.row1 + (n - j.m.rsplit);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int k = kb * 64 + c16 * 8 + i;
        const int scol = k < j.m.ksplit ? j.m.col0 + k : j.m.col1 + (k - j.m.ksplit);
        v[i] = (n < j.rows && k < j.cols) ? j.src[srow * j.ld + scol] : 0.0f;
    }
    uint4 hi, lo;
    if (j.lo16) {
        split2_lo16(v[0], v[1], hi.x, lo.x);
        split2_lo16(v[2], v[3], hi.y, lo.y);
        split2_lo16(v[4], v[5], hi.z, lo.z);
        split2_lo16(v[6], v[7], hi.w, lo.w);
    } else {
        split2(v[0], v[1], hi.x, lo.x);
        split2(v[2], v[3], hi.y, lo.y);
        split2(v[4], v[5], hi.z, lo.z);
        split2(v[6], v[7], hi.w, lo.w);
    }
    size_t base = (size_t)kb * 2 * j.n_pad * 128 + tc::sw128_offset((uint32_t)n, (uint32_t)c16);
    *reinterpret_cast<uint4*>(j.dst + base) = hi;
    *reinterpret_cast<uint4*>(j.dst + base + (size_t)j.n_pad * 128) = lo;
}

static thread_local PackJobs g_jobs;
static thread_local bool g_batching = false;

void pack_batch_begin() {
    g_jobs.n = 0;
    g_batching = true;
}

int pack_batch_flush(cudaStream_t stream) {
    g_batching = false;
    if (g_jobs.n == 0) return HN_OK;
    int max_total = 0;
    for (int i = 0; i < g_jobs.n; ++i) max_total = std::max(max_total, g_jobs.job[i].n_pad * g_jobs.job[i].kblocks * 8);
    pack_b_batch_kernel<<<dim3((unsigned)ceil_div(max_total, 256), (unsigned)g_jobs.n), 256, 0, stream>>>(g_jobs);
    g_jobs.n = 0;
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int launch_pack_b(const float* src, int64_t ld, PackMap map, int rows, int cols, int n_pad, int kblocks, uint8_t* dst,
                  cudaStream_t stream, bool lo16) {
    if (g_batching) {
        if (g_jobs.n == PACK_MAX_JOBS) {            // a full batch: launch it and keep recording
            HN_PROPAGATE(pack_batch_flush(stream));
            g_batching = true;
        }
        g_jobs.job[g_jobs.n++] = PackJob{src, ld, map, rows, cols, n_pad, kblocks, dst, lo16 ? 1 : 0};
        return HN_OK;
    }
    int total = n_pad * kblocks * 8;
    pack_b_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, ld, map, rows, cols, n_pad, kblocks, dst,
                                                                      lo16 ? 1 : 0);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace chain
}  // namespace hn
