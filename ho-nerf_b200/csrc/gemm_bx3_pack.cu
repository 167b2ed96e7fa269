// Packing of a net's weights for gemm_bx3.cuh (HN_TC_BF16X3 on nets without a fused chain kernel: the hand field).
#include "gemm_bx3.cuh"

using namespace hn;

extern "C" {

int64_t hn_mlp_bx3_bytes(const hn_mlp_t* m) {
    if (!m || m->n_layers < 1 || m->n_layers > HN_MAX_LAYERS) return 0;
    return bx3_layout(m).total;
}

int hn_mlp_bx3_pack(const hn_mlp_t* m, void* buf, int64_t bytes, hn_stream_t stream) {
    HN_REQUIRE(m && m->n_layers >= 1 && m->n_layers <= HN_MAX_LAYERS, "hn_mlp_bx3_pack: bad mlp");
    const Bx3Layout L = bx3_layout(m);
    HN_REQUIRE(buf && bytes >= L.total && aligned16(buf), "hn_mlp_bx3_pack: buffer too small or misaligned (need %lld bytes)",
               (long long)L.total);
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* dst = reinterpret_cast<uint8_t*>(buf);
    int jobs = 0;
    chain::pack_batch_begin();
    auto flush_if_full = [&](int need) -> int {
        if (jobs + need > chain::PACK_MAX_JOBS) {
            HN_PROPAGATE(chain::pack_batch_flush(s));
            chain::pack_batch_begin();
            jobs = 0;
        }
        return HN_OK;
    };
    for (int l = 0; l < m->n_layers; ++l) {
        HN_REQUIRE(m->W[l] && m->WT[l], "hn_mlp_bx3_pack: layer %d has no packed weights / transposed copy", l);
        const int out = m->out_dim[l], in = m->in_dim[l];
        const int t_out = (int)ceil_div(out, 256), t_in = (int)ceil_div(in, 256);
        HN_PROPAGATE(flush_if_full(t_out));
        for (int t = 0; t < t_out; ++t, ++jobs)
            HN_PROPAGATE(chain::launch_pack_b(m->W[l], m->ld[l], t * 256, 0, std::min(256, out - t * 256), in, 256,
                                              (int)ceil_div(in, 64), dst + L.w[l] + t * bx3_tile_bytes(in), s));
        HN_PROPAGATE(flush_if_full(t_in));
        for (int t = 0; t < t_in; ++t, ++jobs)
            HN_PROPAGATE(chain::launch_pack_b(m->WT[l], m->ldT[l], t * 256, 0, std::min(256, in - t * 256), out, 256,
                                              (int)ceil_div(out, 64), dst + L.wt[l] + t * bx3_tile_bytes(out), s));
    }
    return chain::pack_batch_flush(s);
}

}  // extern "C"
