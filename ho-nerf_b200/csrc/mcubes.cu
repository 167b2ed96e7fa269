// Marching cubes on the device, after the SDF lattice of extract_geometry (utils/renderer.py:279-283, :561;
// utils/renderer_batch.py:309 call PyMCubes on the host after copying u there; PyMCubes is not part of the reference tree:
// parity unpinned, see ho-nerf_b200/mcubes_tables.py for how the case table is derived).
//
// Mesh with SHARED vertices like PyMCubes': one vertex per lattice edge that crosses the iso value, at the linearly
// interpolated position (index coordinates), triangles as indices into that vertex list.  Four passes, no host round trip
// except the two totals the caller needs to size the outputs:
//   1. hn_mc_classify   per lattice point: which of its three +x / +y / +z edges cross (3 flags);  per cell: its number
//                       of triangles (case table)            -> the caller scans both arrays (torch.cumsum)
//   2. hn_mc_emit       vertices at their scanned slots, triangles of every cell at its scanned offset, each corner
//                       through the scanned slot of the lattice edge it sits on
// HBM-bound integer / byte work: one pass over u per kernel, coalesced along z.
#include <algorithm>

#include "common.cuh"

namespace hn {

struct McTables {
    int8_t n_tris[256];
    int8_t tris[256][15];          // edge ids, up to 5 triangles
    int8_t owner[12][4];           // edge -> (dx, dy, dz, axis) of the lattice point / axis that owns it
};
__constant__ McTables c_mc;

__device__ __forceinline__ int64_t lin(int i, int j, int k, int ny, int nz) { return ((int64_t)i * ny + j) * nz + k; }

// flags[axis][i][j][k] = 1 when the edge from lattice point (i,j,k) along `axis` crosses iso; cell_tris[cell] = #triangles
__global__ void __launch_bounds__(256) mc_classify_kernel(const float* __restrict__ u, int nx, int ny, int nz, float iso,
                                                          int32_t* __restrict__ flags, int32_t* __restrict__ cell_tris) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % nz);
        const int64_t ij = p / nz;
        const int j = (int)(ij % ny), i = (int)(ij / ny);
        const float v = u[p];
        const bool in0 = v < iso;
        flags[p] = (i + 1 < nx) ? (int)(in0 != (u[p + (int64_t)ny * nz] < iso)) : 0;
        flags[n + p] = (j + 1 < ny) ? (int)(in0 != (u[p + nz] < iso)) : 0;
        flags[2 * n + p] = (k + 1 < nz) ? (int)(in0 != (u[p + 1] < iso)) : 0;
        if (i + 1 < nx && j + 1 < ny && k + 1 < nz) {
            // corner c at (i + (c&1 ^ c>>1&1), j + (c>>1&1), k + (c>>2&1)): 0..3 counter-clockwise on z = k, 4..7 above
            const int64_t sx = (int64_t)ny * nz, sy = nz;
            int cs = in0 ? 1 : 0;
            cs |= (u[p + sx] < iso) << 1;
            cs |= (u[p + sx + sy] < iso) << 2;
            cs |= (u[p + sy] < iso) << 3;
            cs |= (u[p + 1] < iso) << 4;
            cs |= (u[p + sx + 1] < iso) << 5;
            cs |= (u[p + sx + sy + 1] < iso) << 6;
            cs |= (u[p + sy + 1] < iso) << 7;
            cell_tris[lin(i, j, k, ny - 1, nz - 1)] = c_mc.n_tris[cs];
        }
    }
}

// vertices[slot] for every flagged edge (slot = inclusive scan - 1); triangles of every cell at its scanned offset
__global__ void __launch_bounds__(256) mc_emit_kernel(const float* __restrict__ u, int nx, int ny, int nz, float iso,
                                                      const int32_t* __restrict__ flags, const int32_t* __restrict__ vscan,
                                                      const int32_t* __restrict__ cell_tris, const int32_t* __restrict__ tscan,
                                                      float* __restrict__ vertices, int32_t* __restrict__ triangles) {
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t sx = (int64_t)ny * nz, sy = nz;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % nz);
        const int64_t ij = p / nz;
        const int j = (int)(ij % ny), i = (int)(ij / ny);
        const float v0 = u[p];
#pragma unroll
        for (int axis = 0; axis < 3; ++axis) {
            if (flags[axis * n + p]) {
                const float v1 = u[p + (axis == 0 ? sx : axis == 1 ? sy : 1)];
                // linear interpolation in double like PyMCubes, rounded once
                const float t = (float)(((double)iso - (double)v0) / ((double)v1 - (double)v0));
                float* o = vertices + 3 * (int64_t)(vscan[axis * n + p] - 1);
                o[0] = (float)i + (axis == 0 ? t : 0.0f);
                o[1] = (float)j + (axis == 1 ? t : 0.0f);
                o[2] = (float)k + (axis == 2 ? t : 0.0f);
            }
        }
        if (i + 1 < nx && j + 1 < ny && k + 1 < nz) {
            const int64_t cell = lin(i, j, k, ny - 1, nz - 1);
            const int nt = cell_tris[cell];
            if (nt > 0) {
                int cs = (v0 < iso) ? 1 : 0;
                cs |= (u[p + sx] < iso) << 1;
                cs |= (u[p + sx + sy] < iso) << 2;
                cs |= (u[p + sy] < iso) << 3;
                cs |= (u[p + 1] < iso) << 4;
                cs |= (u[p + sx + 1] < iso) << 5;
                cs |= (u[p + sx + sy + 1] < iso) << 6;
                cs |= (u[p + sy + 1] < iso) << 7;
                int32_t* o = triangles + 3 * (int64_t)(tscan[cell] - nt);
                for (int q = 0; q < 3 * nt; ++q) {
                    const int e = c_mc.tris[cs][q];
                    const int8_t* ow = c_mc.owner[e];
                    const int64_t pe = p + ow[0] * sx + ow[1] * sy + ow[2];
                    o[q] = vscan[ow[3] * n + pe] - 1;
                }
            }
        }
    }
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_mc_set_tables(const int8_t* n_tris, const int8_t* tris, const int8_t* owner) {
    HN_REQUIRE(n_tris && tris && owner, "hn_mc_set_tables: null argument");
    McTables t;
    for (int c = 0; c < 256; ++c) {
        t.n_tris[c] = n_tris[c];
        for (int q = 0; q < 15; ++q) t.tris[c][q] = tris[c * 15 + q];
    }
    for (int e = 0; e < 12; ++e)
        for (int q = 0; q < 4; ++q) t.owner[e][q] = owner[e * 4 + q];
    HN_CHECK_CUDA(cudaMemcpyToSymbol(c_mc, &t, sizeof(t)));
    return HN_OK;
}

int hn_mc_classify(const float* u, int nx, int ny, int nz, float iso, int32_t* flags, int32_t* cell_tris, hn_stream_t stream) {
    HN_REQUIRE(u && flags && cell_tris && nx >= 2 && ny >= 2 && nz >= 2, "hn_mc_classify: bad arguments");
    const int64_t n = (int64_t)nx * ny * nz;
    HN_REQUIRE(3 * n < ((int64_t)1 << 31), "hn_mc_classify: lattice too large for 32-bit slots");
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
    mc_classify_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(u, nx, ny, nz, iso, flags, cell_tris);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_mc_emit(const float* u, int nx, int ny, int nz, float iso, const int32_t* flags, const int32_t* vscan,
               const int32_t* cell_tris, const int32_t* tscan, float* vertices, int32_t* triangles, hn_stream_t stream) {
    HN_REQUIRE(u && flags && vscan && cell_tris && tscan && nx >= 2 && ny >= 2 && nz >= 2, "hn_mc_emit: bad arguments");
    const int64_t n = (int64_t)nx * ny * nz;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
    mc_emit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(u, nx, ny, nz, iso, flags, vscan, cell_tris, tscan, vertices, triangles);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
