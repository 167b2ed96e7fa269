// Job descriptions of the weight-gradient kernel (chain_dw.cu).
#pragma once
#include <stdint.h>

namespace hn {
namespace chain {

constexpr int DW_MAX_JOBS = 12;
constexpr int DW_SPLITS = 16;

// fp32 operand [points, cols]: row-major (tiled = 0), the chain stash layout [tile][col/4][128 rows][4] (tiled = 1), or
// column-major tiles [tile][ld columns][128 rows] (tiled = 2, the 64-wide encoding arrays)
struct DwOperand {
    const float* ptr;
    int64_t ld;        // row-major leading dimension / columns per column-major tile (ignored when tiled = 1)
    int cols;          // valid columns (the rest of the 256-wide tile is zero)
    int tiled;
};
struct DwJob {
    DwOperand P[2], Q[2];
    int n_pairs;
    int n_mma;         // UMMA N = round_up(Q cols, 16)
    float* db;         // += column sums of P[0] (may be NULL)
    float db_scale;
};
struct DwParams {
    int64_t n;
    int n_tiles;
    int n_jobs;
    DwJob job[DW_MAX_JOBS];
    float* part;       // [n_jobs][DW_SPLITS][256][256]
};

// dW[(row0 + r) * ld + col(c)] += sum_s part[job][s][r][c]   for r < rows, c < cols, where
// col(c) = c < csplit ? col0 + c : col1 + (c - csplit)
struct DwReduceJob {
    float* dW;
    int ld, row0, rows, cols;
    int col0, csplit, col1;
};
inline DwReduceJob reduce_job(float* dW, int ld, int row0, int rows, int cols, int col0 = 0, int csplit = 1 << 30, int col1 = 0) {
    return DwReduceJob{dW, ld, row0, rows, cols, col0, csplit, col1};
}
struct DwReduceParams {
    DwReduceJob job[DW_MAX_JOBS];
    const float* part;
};

int64_t dw_part_floats(int n_jobs);
int launch_dw(const DwParams& p, const DwReduceParams& r, cudaStream_t s);
int launch_dw_reduce(const DwReduceParams& r, int n_jobs, cudaStream_t s);

}  // namespace chain
}  // namespace hn
