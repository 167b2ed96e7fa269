// Ray generation: NDC pixel coordinates -> world-space ray origins and unit directions,
// replacing utils/utils.py:31-115 (_xy_to_ray_bundle) + pytorch3d PerspectiveCameras.unproject_points(from_ndc=True)
// and, for full-image renders, the NDC grid of exp_runner.py:338-350, so a render never materialises the
// [2 n, 3] un-projected planes (and, for the grid form, not even the xy list).  One thread per ray: 8 B in, 24 B out.
//
// pytorch3d convention restated (row vectors): X_view = X_world R + T;  x_ndc = fx X/Z + px,  y_ndc = fy Y/Z + py.
// Un-projecting (x, y) at depth z gives X_view = ((x - px) z / fx, (y - py) z / fy, z),  X_world = (X_view - T) R^-1.
// The bundle is built from the depth-1 and depth-2 planes:  d = normalize(P2 - P1),  o = P1 - d  (utils/utils.py:96-107).
#include "common.cuh"

namespace hn {

struct Camera {
    float Rinv[9];   // R^-1, row-major
    float T[3];
    float inv_fx, inv_fy, px, py;
};

// cam: 16 floats  R[9] (row-major) | T[3] | fx fy | px py
__device__ __forceinline__ Camera load_camera(const float* __restrict__ cam) {
    Camera c;
    float r[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = cam[k];
    // inverse through the adjugate: exact transpose for an orthonormal R up to rounding, still right for any regular R
    float c00 = r[4] * r[8] - r[5] * r[7], c01 = r[5] * r[6] - r[3] * r[8], c02 = r[3] * r[7] - r[4] * r[6];
    float det = r[0] * c00 + r[1] * c01 + r[2] * c02;
    float id = 1.0f / det;
    c.Rinv[0] = c00 * id;
    c.Rinv[1] = (r[2] * r[7] - r[1] * r[8]) * id;
    c.Rinv[2] = (r[1] * r[5] - r[2] * r[4]) * id;
    c.Rinv[3] = c01 * id;
    c.Rinv[4] = (r[0] * r[8] - r[2] * r[6]) * id;
    c.Rinv[5] = (r[2] * r[3] - r[0] * r[5]) * id;
    c.Rinv[6] = c02 * id;
    c.Rinv[7] = (r[1] * r[6] - r[0] * r[7]) * id;
    c.Rinv[8] = (r[0] * r[4] - r[1] * r[3]) * id;
#pragma unroll
    for (int k = 0; k < 3; ++k) c.T[k] = cam[9 + k];
    c.inv_fx = 1.0f / cam[12];
    c.inv_fy = 1.0f / cam[13];
    c.px = cam[14];
    c.py = cam[15];
    return c;
}

__device__ __forceinline__ void unproject(const Camera& c, float x, float y, float depth, float (&w)[3]) {
    float v[3] = {(x - c.px) * c.inv_fx * depth - c.T[0], (y - c.py) * c.inv_fy * depth - c.T[1], depth - c.T[2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = v[0] * c.Rinv[k] + v[1] * c.Rinv[3 + k] + v[2] * c.Rinv[6 + k];
}

__device__ __noinline__ void make_ray(const Camera& c, float x, float y, float* __restrict__ o, float* __restrict__ d) {
    float p1[3], p2[3], dir[3];
    unproject(c, x, y, 1.0f, p1);
    unproject(c, x, y, 2.0f, p2);
#pragma unroll
    for (int k = 0; k < 3; ++k) dir[k] = p2[k] - p1[k];
    float nrm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float inv = 1.0f / fmaxf(nrm, 1e-12f);                                 // F.normalize: x / max(|x|, eps)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float u = dir[k] * inv;
        d[k] = u;
        o[k] = p1[k] - u;
    }
}

__global__ void rays_from_ndc_kernel(const float* __restrict__ xy, const float* __restrict__ cams, int64_t n_per_cam,
                                     int64_t total, float* __restrict__ rays_o, float* __restrict__ rays_d) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Camera c = load_camera(cams + (i / n_per_cam) * 16);
    float2 p = reinterpret_cast<const float2*>(xy)[i];
    make_ray(c, p.x, p.y, rays_o + i * 3, rays_d + i * 3);
}

__global__ void rays_ndc_grid_kernel(const float* __restrict__ xs, const float* __restrict__ ys, int W,
                                     const float* __restrict__ cam, int64_t first, int64_t count,
                                     float* __restrict__ rays_o, float* __restrict__ rays_d) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Camera c = load_camera(cam);
    int64_t pix = first + i;
    int64_t row = pix / W;
    int col = (int)(pix - row * W);
    make_ray(c, xs[col], ys[row], rays_o + i * 3, rays_d + i * 3);
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_rays_from_ndc(const float* xy, const float* cams, int64_t n_cams, int64_t n_per_cam, float* rays_o,
                     float* rays_d, hn_stream_t stream) {
    HN_REQUIRE(n_cams >= 0 && n_per_cam >= 0, "hn_rays_from_ndc: bad sizes");
    int64_t total = n_cams * n_per_cam;
    if (total == 0) return HN_OK;
    HN_REQUIRE(xy && cams && rays_o && rays_d, "hn_rays_from_ndc: null pointer");
    HN_REQUIRE((reinterpret_cast<uintptr_t>(xy) & 7) == 0, "hn_rays_from_ndc: xy must be 8-byte aligned");
    rays_from_ndc_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(xy, cams, n_per_cam, total,
                                                                                            rays_o, rays_d);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_rays_ndc_grid(const float* xs, const float* ys, int W, int H, const float* cam, int64_t first, int64_t count,
                     float* rays_o, float* rays_d, hn_stream_t stream) {
    HN_REQUIRE(W > 0 && H > 0 && first >= 0 && count >= 0 && first + count <= (int64_t)W * H,
               "hn_rays_ndc_grid: pixel range [%lld, %lld) outside the %d x %d image", (long long)first,
               (long long)(first + count), W, H);
    if (count == 0) return HN_OK;
    HN_REQUIRE(xs && ys && cam && rays_o && rays_d, "hn_rays_ndc_grid: null pointer");
    rays_ndc_grid_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, (cudaStream_t)stream>>>(xs, ys, W, cam, first, count,
                                                                                            rays_o, rays_d);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
