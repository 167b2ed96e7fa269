// Nearest-neighbour selection of the temporal contact-stability loss, replacing the
// per-frame scipy cKDTree round trips of utils/renderer_batch.py:346-357: for every frame t and every "in" point p
// (in_mask[t,p] != 0) find the nearest point q among the frame's "out" candidates (out_mask[t,q] != 0) and raise
// flag[t,q] -- np.unique(near_out_id) as a flag array.  Brute force, one warp per (t, p): the clouds are object
// vertices / 10 (hundreds to a few thousand points), so T * P^2 distance evaluations are microseconds and nothing
// leaves the device.  Distances are evaluated in fp64 from the fp32 coordinates, like cKDTree does, so the selected
// index is the one scipy returns whenever the minimum is unique; exact ties go to the lowest index.
#include "common.cuh"

namespace hn {

__global__ void __launch_bounds__(256)
nn_select_kernel(const float* __restrict__ pts, const uint8_t* __restrict__ in_mask,
                 const uint8_t* __restrict__ out_mask, int T, int P, uint8_t* __restrict__ flag,
                 int64_t* __restrict__ nearest) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= (int64_t)T * P) return;
    int t = (int)(w / P), p = (int)(w - (int64_t)t * P);
    if (!in_mask[w]) {
        if (nearest && lane == 0) nearest[w] = -1;
        return;
    }
    double px = pts[p * 3 + 0], py = pts[p * 3 + 1], pz = pts[p * 3 + 2];
    const uint8_t* om = out_mask + (int64_t)t * P;
    double best = 1.0e300;
    int best_q = 0x7fffffff;
    for (int q = lane; q < P; q += 32) {
        if (!om[q]) continue;
        double dx = (double)pts[q * 3 + 0] - px, dy = (double)pts[q * 3 + 1] - py, dz = (double)pts[q * 3 + 2] - pz;
        double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) { best = d2; best_q = q; }          // q ascending per lane: ties keep the lowest index
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oq = __shfl_xor_sync(0xffffffffu, best_q, o);
        if (ob < best || (ob == best && oq < best_q)) { best = ob; best_q = oq; }
    }
    if (lane == 0) {
        bool found = best_q != 0x7fffffff;
        if (found) flag[(int64_t)t * P + best_q] = 1;      // racing writers all store 1
        if (nearest) nearest[w] = found ? best_q : -1;
    }
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_nn_select(const float* pts, const uint8_t* in_mask, const uint8_t* out_mask, int n_frames, int n_pts,
                 uint8_t* flag, int64_t* nearest, hn_stream_t stream) {
    HN_REQUIRE(n_frames >= 0 && n_pts >= 0, "hn_nn_select: bad sizes");
    int64_t warps = (int64_t)n_frames * n_pts;
    if (warps == 0) return HN_OK;
    HN_REQUIRE(pts && in_mask && out_mask && flag, "hn_nn_select: null pointer");
    HN_REQUIRE(warps * 32 / 256 < 0x7fffffff, "hn_nn_select: %d frames x %d points is too large", n_frames, n_pts);
    nn_select_kernel<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(pts, in_mask, out_mask,
                                                                                            n_frames, n_pts, flag, nearest);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
