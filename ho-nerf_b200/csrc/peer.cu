// The one exchange step of the path (the flat gradient of a ray-sharded training step), fused with the
// optimiser: gradient all-reduce + Adam as ONE kernel over NVLink peer memory (exp_runner.py:83 builds one Adam over every
// parameter; the reference is single-GPU, so the collective has no upstream counterpart).
//
// Every rank owns one peer-mapped block  [ g : n floats | red : n floats | sig : PEER_MAX_CTAS * world words ]  (cudaMalloc +
// cudaIpcGetMemHandle, opened by the other ranks with cudaIpcOpenMemHandle: NVLink / NVSwitch loads and stores).
// Two-shot, per-CTA pipelines -- CTA b of every rank works with CTA b of every other rank and with nobody else:
//   barrier 1 (all ranks' gradients are complete)
//   phase 1   rank r sums chunk (r, b) of the gradient over all ranks, IN RANK ORDER, into its own `red`
//   barrier 2
//   phase 2   every rank reads the reduced chunks (0..world-1, b) from their owners and applies Adam to those elements
// Each element is summed once, by one rank, and broadcast: the parameters stay bit-identical across ranks.  Per GPU
// 2 (world-1)/world n floats cross NVLink (a one-shot sum would read (world-1) n).  `red` is rewritten only after the NEXT
// launch's barrier 1, which every rank reaches after its previous launch completed: no trailing barrier is needed.
// A barrier is a flag per (CTA, source rank) holding a monotonically increasing epoch: st.release.sys into every peer's
// flag array, ld.acquire.sys spins on the own one.  The epoch lives in device memory (per CTA), so a captured CUDA graph
// replays correctly.  A spin that outlasts PEER_TIMEOUT_NS (60 s) sets *err and lets the CTA run on (garbage, but no hung GPU):
// the host checks the word (hn_peer_adam_flat's caller: optim.FlatAdam.peer_error()).
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace hn {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_MAX_CTAS = 148;
constexpr int PEER_THREADS = 1024;
constexpr unsigned long long PEER_TIMEOUT_NS = 60000000000ull;

struct PeerPtrs {
    const float* g[PEER_MAX_WORLD];
    float* red[PEER_MAX_WORLD];
    uint32_t* sig[PEER_MAX_WORLD];
};

__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// CTA b of every rank meets here.  Threads 0 .. world-1 each talk to one peer.
__device__ __forceinline__ void peer_barrier(const PeerPtrs& P, int rank, int world, uint32_t target, uint32_t* err) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < world) {
        const int slot = blockIdx.x * world;
        st_release_sys(P.sig[t] + slot + rank, target);
        const uint32_t* mine = P.sig[rank] + slot + t;
        const unsigned long long t0 = global_ns();
        int polls = 0;
        while ((int32_t)(ld_acquire_sys(mine) - target) < 0) {
            if ((++polls & 1023) == 0 && global_ns() - t0 > PEER_TIMEOUT_NS) {
                atomicExch(err, 1u + (uint32_t)t);
                break;
            }
        }
    }
    __syncthreads();
}

// chunk (r, b): float4 range of rank r's slice handled by CTA b
__device__ __forceinline__ void peer_chunk(int64_t q_total, int world, int r, int b, int n_ctas, int64_t& lo, int64_t& hi) {
    const int64_t s_lo = q_total * r / world, s_hi = q_total * (r + 1) / world;
    lo = s_lo + (s_hi - s_lo) * b / n_ctas;
    hi = s_lo + (s_hi - s_lo) * (b + 1) / n_ctas;
}

// mode 0: Adam on (p, m, v);  mode 1: p[i] = grad_scale * summed gradient (the plain all-reduce, used to check the exchange)
// Both phases are LATENCY-bound (an NVLink load takes ~3 us; measured: the first version, one load per thread and loop trip,
// took 10 us + 3 us per trip): every thread issues all the remote loads of its elements before it uses any of them.
template <int MODE>
__global__ void __launch_bounds__(PEER_THREADS) peer_adam_kernel(PeerPtrs P, int rank, int world, float* __restrict__ p,
                                                                 float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                                 uint32_t* __restrict__ epoch, uint32_t* __restrict__ err,
                                                                 const float* __restrict__ step, const float* __restrict__ lr_dev,
                                                                 float lr, float beta1, float beta2, float omb1, float omb2, float eps,
                                                                 float weight_decay, float grad_scale) {
    __shared__ int64_t c_lo[PEER_MAX_WORLD];          // phase 2: chunk (r, b) starts at float4 c_lo[r] ...
    __shared__ int c_pre[PEER_MAX_WORLD + 1];         // ... and holds elements [c_pre[r], c_pre[r + 1]) of this CTA's work list
    const int b = blockIdx.x, n_ctas = gridDim.x;
    const uint32_t e0 = epoch[b];
    const int64_t q_total = n >> 2;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int rr = 0; rr < world; ++rr) {
            const int r = (rank + rr) % world;          // the own chunk first: the ranks spread over the owners
            int64_t lo, hi;
            peer_chunk(q_total, world, r, b, n_ctas, lo, hi);
            c_lo[rr] = lo;
            c_pre[rr] = acc;
            acc += (int)(hi - lo);
        }
        c_pre[world] = acc;
    }
    peer_barrier(P, rank, world, e0 + 1u, err);
    {
        int64_t lo, hi;
        peer_chunk(q_total, world, rank, b, n_ctas, lo, hi);
        float4* out = reinterpret_cast<float4*>(P.red[rank]);
        for (int64_t i0 = lo + threadIdx.x; i0 < hi; i0 += 2 * PEER_THREADS) {
            const int64_t i1 = i0 + PEER_THREADS;
            const bool two = i1 < hi;
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            for (int r0 = 0; r0 < world; r0 += 4) {       // rank order 0, 1, 2, ...: every element has ONE summation order
                float4 x0[4], x1[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (r0 + k < world) {
                        x0[k] = ld_sys_f4(P.g[r0 + k] + 4 * i0);
                        if (two) x1[k] = ld_sys_f4(P.g[r0 + k] + 4 * i1);
                    }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (r0 + k < world) {
                        if (r0 + k == 0) {
                            s0 = x0[k];
                            if (two) s1 = x1[k];
                        } else {
                            s0.x += x0[k].x; s0.y += x0[k].y; s0.z += x0[k].z; s0.w += x0[k].w;
                            if (two) { s1.x += x1[k].x; s1.y += x1[k].y; s1.z += x1[k].z; s1.w += x1[k].w; }
                        }
                    }
            }
            out[i0] = s0;
            if (two) out[i1] = s1;
        }
    }
    peer_barrier(P, rank, world, e0 + 2u, err);
    float step_size = 0.0f, inv_sqrt_bc2 = 0.0f;
    if (MODE == 0) {
        const float t0 = *step;
        if (lr_dev) lr = *lr_dev;
        const float bc1 = 1.0f - powf(beta1, t0), bc2 = 1.0f - powf(beta2, t0);
        step_size = lr / bc1;
        inv_sqrt_bc2 = rsqrtf(bc2);
    }
    const int total = c_pre[world];
    constexpr int E = 2;
    for (int e_base = 0; e_base < total; e_base += E * PEER_THREADS) {
        int64_t idx[E];
        float4 g4[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int e = e_base + k * PEER_THREADS + threadIdx.x;
            idx[k] = -1;
            if (e < total) {
                int rr = 0;
                while (e >= c_pre[rr + 1]) ++rr;
                idx[k] = c_lo[rr] + (e - c_pre[rr]);
                g4[k] = ld_sys_f4(P.red[(rank + rr) % world] + 4 * idx[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < E; ++k) {
            if (idx[k] < 0) continue;
            const int64_t i = idx[k];
            const float gs[4] = {g4[k].x * grad_scale, g4[k].y * grad_scale, g4[k].z * grad_scale, g4[k].w * grad_scale};
            float4* p4 = reinterpret_cast<float4*>(p) + i;
            if (MODE == 1) {
                *p4 = make_float4(gs[0], gs[1], gs[2], gs[3]);
                continue;
            }
            float4* m4 = reinterpret_cast<float4*>(m) + i;
            float4* v4 = reinterpret_cast<float4*>(v) + i;
            float4 pv = *p4, mv = *m4, vv = *v4;
            float* pp = &pv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float gi = gs[c];
                if (weight_decay != 0.0f) gi += weight_decay * pp[c];
                const float mi = beta1 * mm[c] + omb1 * gi;
                const float vi = beta2 * vq[c] + omb2 * gi * gi;
                mm[c] = mi;
                vq[c] = vi;
                pp[c] = pp[c] - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
            }
            *p4 = pv; *m4 = mv; *v4 = vv;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) epoch[b] = e0 + 2u;
}

}  // namespace hn

using namespace hn;

extern "C" int64_t hn_peer_block_bytes(int64_t n, int world) {
    if (n < 0 || world < 1 || world > PEER_MAX_WORLD) return -1;
    return 2 * round_up(n, 64) * 4 + (int64_t)PEER_MAX_CTAS * world * 4 + 256;
}

extern "C" int hn_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle64) {
    HN_REQUIRE(ptr && handle64 && bytes > 0, "hn_peer_alloc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    HN_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
    HN_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
    HN_CHECK_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("hn_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return HN_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return HN_OK;
}

extern "C" int hn_peer_open(const uint8_t* handle64, void** ptr) {
    HN_REQUIRE(ptr && handle64, "hn_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    HN_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return HN_OK;
}

extern "C" int hn_peer_close(void* ptr) {
    if (ptr) HN_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return HN_OK;
}

extern "C" int hn_peer_free(void* ptr) {
    if (ptr) HN_CHECK_CUDA(cudaFree(ptr));
    return HN_OK;
}

extern "C" int hn_peer_adam_flat(float* p, float* m, float* v, int64_t n, const void* const* blocks, int rank, int world,
                                 uint32_t* epoch, uint32_t* err, int mode, const float* step, const float* lr_dev, double lr,
                                 double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                                 hn_stream_t stream) {
    HN_REQUIRE(p && blocks && epoch && err && n > 0, "hn_peer_adam_flat: null argument");
    HN_REQUIRE(mode == 1 || (m && v && step), "hn_peer_adam_flat: Adam state missing");
    HN_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "hn_peer_adam_flat: rank %d of %d", rank, world);
    HN_REQUIRE((n & 3) == 0 && aligned16(p) && (mode == 1 || (aligned16(m) && aligned16(v))),
               "hn_peer_adam_flat: n must be a multiple of 4 and the buffers 16-byte aligned");
    PeerPtrs P;
    const int64_t stride = round_up(n, 64) * 4;
    for (int r = 0; r < world; ++r) {
        HN_REQUIRE(blocks[r], "hn_peer_adam_flat: block of rank %d is null", r);
        uint8_t* base = (uint8_t*)blocks[r];
        P.g[r] = (const float*)base;
        P.red[r] = (float*)(base + stride);
        P.sig[r] = (uint32_t*)(base + 2 * stride);
    }
    for (int r = world; r < PEER_MAX_WORLD; ++r) { P.g[r] = nullptr; P.red[r] = nullptr; P.sig[r] = nullptr; }
    int max_ctas = std::min(PEER_MAX_CTAS, sm_count());
    if (const char* e = getenv("HONERF_PEER_CTAS")) {           // tuning aid (tools/prof_peer.py); must be equal on every rank
        const int v = atoi(e);
        if (v >= 1 && v < max_ctas) max_ctas = v;
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(max_ctas, (n >> 2) / (64 * world) + 1));
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 1)
        peer_adam_kernel<1><<<grid, PEER_THREADS, 0, s>>>(P, rank, world, p, m, v, n, epoch, err, step, lr_dev, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
                                                          0.f, (float)grad_scale);
    else
        peer_adam_kernel<0><<<grid, PEER_THREADS, 0, s>>>(P, rank, world, p, m, v, n, epoch, err, step, lr_dev, (float)lr, (float)beta1,
                                                          (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                          (float)weight_decay, (float)grad_scale);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}
