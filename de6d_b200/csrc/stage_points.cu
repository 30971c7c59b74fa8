// Input staging for sm_100a: the step right in front of the sampling path.
//
// The reference gets from "a list of (n_i, 3+C) point arrays" to the two tensors the SA layers read in five host/device
// passes: DataProcessor.sample_points gathers points[choice] on the host (data_processor.py:145-177), collate pads a
// batch-index column and concatenates, load_data_to_gpu copies (models/__init__.py:23-34), and
// PointNet2FSMSG.forward (pointnet2_backbone.py:193-222) slices xyz / features out of the (B*N, 4+C) array, counts the
// rows of every frame with one host-synchronising .sum() per frame, and makes a contiguous (B,N,3) and a permuted
// contiguous (B,C,N) copy.
//
// Here one kernel reads each source row once and writes xyz (B,N,3), features (B,C,N) and (optionally) the float
// batch index (B,N) directly.  `choice` (optional) is the host-drawn sample_points index list, already offset to global
// rows, so the gather happens in the same pass; `lead` says whether the rows carry the collated batch-index column.
// The frame-size check of pointnet2_backbone.py:214-218 becomes a device-side counter (status[0] = rows whose batch
// column differs from the frame slot they land in), read back once by the caller instead of B times.
//
// Pure copies: bit-exact by construction.  HBM-bound: 4*(lead+3+C) B read + 4*(3+C) B written per point.
#include "common.cuh"

namespace de6d {

constexpr int STAGE_THREADS = 256;

template <bool LEAD>
__global__ void __launch_bounds__(STAGE_THREADS)
stage_points_kernel(int n, int c, long long total_rows, const float *__restrict__ src, const int *__restrict__ choice,
                    float *__restrict__ xyz, float *__restrict__ features, float *__restrict__ batch_idx,
                    int *__restrict__ status) {
    __shared__ float sx[STAGE_THREADS * 3];
    const int bs = blockIdx.y;
    const int i0 = blockIdx.x * STAGE_THREADS;
    const int i = i0 + threadIdx.x;
    const int w = (LEAD ? 4 : 3) + c;
    const bool live = i < n;
    long long row = -1;
    if (live) row = choice ? (long long)__ldg(choice + (size_t)bs * n + i) : (long long)bs * n + i;
    const bool ok = live && row >= 0 && row < total_rows;
    const float *r = src + (size_t)(ok ? row : 0) * w;
    float x = 0.f, y = 0.f, z = 0.f, bi = (float)bs;
    if (ok) {
        if (LEAD) bi = r[0];
        x = r[LEAD]; y = r[LEAD + 1]; z = r[LEAD + 2];
    }
    sx[threadIdx.x * 3] = x; sx[threadIdx.x * 3 + 1] = y; sx[threadIdx.x * 3 + 2] = z;
    if (features && live) {
        float *f = features + (size_t)bs * c * n + i;
        const float *rf = r + (LEAD ? 4 : 3);
        for (int ch = 0; ch < c; ++ch) f[(size_t)ch * n] = ok ? rf[ch] : 0.f;
    }
    if (batch_idx && live) batch_idx[(size_t)bs * n + i] = bi;
    if (status) {
        const unsigned bad_b = __ballot_sync(0xffffffffu, ok && LEAD && bi != (float)bs);
        const unsigned bad_r = __ballot_sync(0xffffffffu, live && !ok);
        if ((threadIdx.x & 31) == 0) {
            if (bad_b) atomicAdd(status, __popc(bad_b));
            if (bad_r) atomicAdd(status + 1, __popc(bad_r));
        }
    }
    __syncthreads();
    // the CTA's xyz rows are contiguous in the output: write them back coalesced
    const int cnt = 3 * min(STAGE_THREADS, n - i0);
    float *o = xyz + ((size_t)bs * n + i0) * 3;
    for (int k = threadIdx.x; k < cnt; k += STAGE_THREADS) o[k] = sx[k];
}

}  // namespace de6d

using namespace de6d;

// src (total_rows, lead+3+c) f32 rows [batch_idx?, x, y, z, feat...]; choice NULL (row = bs*n+i; needs total_rows == b*n)
// or (b,n) i32 global row indices; xyz (b,n,3); features (b,c,n) or NULL when c == 0; batch_idx (b,n) f32 or NULL;
// status i32[2] or NULL, zeroed here: [0] rows whose batch column != frame slot (lead only), [1] choice entries outside
// [0,total_rows) (written as zeros).
extern "C" int de6d_stage_points(int b, int n, int c, int lead, long long total_rows, const float *src, const int *choice,
                                 float *xyz, float *features, float *batch_idx, int *status, cudaStream_t stream) {
    if (b < 0 || n < 0 || c < 0 || total_rows < 0) return de6d_set_error(DE6D_ERR_INVALID, "stage_points: negative size");
    if (lead != 0 && lead != 1) return de6d_set_error(DE6D_ERR_INVALID, "stage_points: lead must be 0 or 1");
    if (!choice && total_rows != (long long)b * n)
        return de6d_set_error(DE6D_ERR_INVALID, "stage_points: without choice the source must hold exactly b*n rows");
    if (status) {
        cudaError_t e = cudaMemsetAsync(status, 0, 2 * sizeof(int), stream);
        if (e != cudaSuccess) return de6d_set_error(DE6D_ERR_CUDA, cudaGetErrorString(e));
    }
    if (b == 0 || n == 0) return DE6D_OK;
    if (!xyz || (total_rows > 0 && !src) || (c > 0 && !features))
        return de6d_set_error(DE6D_ERR_INVALID, "stage_points: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "stage_points: batch > 65535");
    dim3 grid(ceil_div(n, STAGE_THREADS), b);
    if (lead)
        stage_points_kernel<true><<<grid, STAGE_THREADS, 0, stream>>>(n, c, total_rows, src, choice, xyz,
                                                                      c > 0 ? features : nullptr, batch_idx, status);
    else
        stage_points_kernel<false><<<grid, STAGE_THREADS, 0, stream>>>(n, c, total_rows, src, choice, xyz,
                                                                       c > 0 ? features : nullptr, batch_idx, status);
    DE6D_CHECK_LAUNCH("stage_points_kernel");
    return DE6D_OK;
}
