// D-FPS / S-FPS of small clouds (32 <= N <= 4096) for sm_100a: register-resident, no pruning, (almost) no barriers.
//
// The bucket kernel of fps.cu pays ~900 cycles of fixed work per selected point (lower bounds, visit dispatch, two
// arg-max levels, a CTA barrier) however small the cloud is: 512 -> 256 took as long per sample as 4096 -> 512.  For a
// small cloud it is cheaper to touch every point every time, provided nothing leaves the registers:
//   * one warp per cloud up to 1024 points (four clouds per CTA, no barrier at all in the loop), four warps per cloud
//     up to 4096 points (one barrier per sample); every lane keeps x, y, z, min-distance (and weight) of its <= 32
//     points in registers;
//   * the points of a lane are arranged so that its register order IS the reference's tie order: point
//     k = c*B + (bitrev(a) << 5) + lane sits in slot J = a*C + c (B = opt_n_threads(N), C = ceil(N / B)), whose tie
//     priority (common.cuh: fps_prio) is  bitrev5(lane) : a : c  -- ascending in J.  The in-lane arg-max is then a
//     plain strict '>' scan in register order (one compare + two selects per point), exactly the reference's in-thread
//     rule (sampling_gpu.cu:146-147); lanes / warps are merged with REDUX on (order-preserving value, priority);
//   * the winner's coordinates come from a shared-memory copy of the cloud indexed by the winning point.
// Same arithmetic as the reference (common.cuh: sqdist, fps.cu: sfps_key), bit-identical indices and min-distances.
#include "common.cuh"
#include <math.h>
#include <type_traits>

namespace de6d {

__device__ __forceinline__ float small_sfps_key(float d, float w) {   // see fps.cu: sfps_key
    if (w >= 1e-11f) return __fmul_rn(d, w);
    return (float)((double)d * fmax((double)w, 1e-12));
}

constexpr int FS_THREADS = 128;

// W warps per cloud (1 or 4), RPL points per lane.  CTA = 4 warps = 4 / W clouds.
template <int MODE, int W, int RPL>
__global__ void __launch_bounds__(FS_THREADS)
fps_small_kernel(int b, int n, int m, int log2B, int C, int cinv, const float *__restrict__ xyz_all,
                 const float *__restrict__ w_all, float *__restrict__ temp_all, int *__restrict__ idx_all) {
    extern __shared__ __align__(16) float fs_smem[];
    __shared__ uint2 wbuf[2][4];
    constexpr int CPB = 4 / W;                      // clouds per CTA
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ww = warp % W;                        // warp index inside its cloud
    const int cloud = blockIdx.x * CPB + warp / W;
    if (cloud >= b) return;   // W == 4: whole CTAs; W == 1: whole warps, and a one-warp cloud uses no CTA barrier below
    const float *xyz = xyz_all + (size_t)cloud * n * 3;
    const float *wts = MODE == 1 ? w_all + (size_t)cloud * n : nullptr;
    float *temp_g = temp_all + (size_t)cloud * n;
    int *idxs = idx_all + (size_t)cloud * m;
    float *sxyz = fs_smem + (size_t)(warp / W) * n * 3;   // this cloud's coordinates, AoS like the input

    const int hbits = log2B - 5;
    const uint32_t B = 1u << log2B;
    const uint32_t lanepart = (__brev((uint32_t)lane) >> 27) << (hbits + 22);

    // ---- load: coordinates -> smem copy + registers in tie order --------------------------------------------------
    if (W == 1) for (int e = lane; e < 3 * n; e += 32) sxyz[e] = xyz[e];
    else for (int e = tid; e < 3 * n; e += FS_THREADS) sxyz[e] = xyz[e];
    float x[RPL], y[RPL], z[RPL], t[RPL], wt[MODE == 1 ? RPL : 1];
#pragma unroll
    for (int jj = 0; jj < RPL; ++jj) {
        const uint32_t J = (uint32_t)(ww * RPL + jj);
        const uint32_t a = J / (uint32_t)C, c = J - a * (uint32_t)C;
        const uint32_t k = c * B + ((hbits ? (__brev(a) >> (32 - hbits)) : 0u) << 5) + (uint32_t)lane;
        const bool valid = a < (B >> 5) && k < (uint32_t)n;
        x[jj] = valid ? xyz[k * 3 + 0] : 0.f;
        y[jj] = valid ? xyz[k * 3 + 1] : 0.f;
        z[jj] = valid ? xyz[k * 3 + 2] : 0.f;
        t[jj] = valid ? temp_g[k] : -INFINITY;       // fminf(d, -inf) stays -inf: padding never wins, never changes
        if (MODE == 1) wt[jj] = valid ? wts[k] : 0.f;
    }
    if (W == 1) __syncwarp(); else __syncthreads();

    constexpr uint32_t ORD_M1 = 0x407fffffu;         // f2ord(-1.0f): a value must exceed -1 to be selected
    int par = 0;
    // merge the lanes' (best key, slot) into the cloud's winner; returns the selected point index
    auto select = [&](float bt, int bj) -> int {
        const uint32_t J = (uint32_t)(ww * RPL + bj);
        const uint32_t a = (J * (uint32_t)cinv) >> 16, c = J - a * (uint32_t)C;
        uint32_t v = f2ord(bt), prio = lanepart | (a << 22) | c;    // bt is never NaN (NaN fails '>')
        warp_argmax(v, prio);
        if (W > 1) {
            if (lane == 0) wbuf[par][ww] = make_uint2(v, prio);
            __syncthreads();
            uint2 best = wbuf[par][0];    // W entries, read by every lane: a compare chain is shorter than two REDUX
#pragma unroll
            for (int i = 1; i < W; ++i) {
                const uint2 e = wbuf[par][i];
                if (e.x > best.x || (e.x == best.x && e.y < best.y)) best = e;
            }
            par ^= 1;
            v = best.x; prio = best.y;
        }
        return v > ORD_M1 ? (int)fps_prio_to_index(prio, (uint32_t)log2B) : 0;
    };

    int old = 0, first_it = 1;
    if (MODE == 1) {   // S-FPS starts from argmax(weights) with the same candidate rule (sampling_gpu.cu:451-455)
        float bt = -INFINITY;
        int bj = 0;
#pragma unroll
        for (int jj = 0; jj < RPL; ++jj) {
            const uint32_t J = (uint32_t)(ww * RPL + jj);
            const uint32_t a = J / (uint32_t)C, c = J - a * (uint32_t)C;
            const uint32_t k = c * B + ((hbits ? (__brev(a) >> (32 - hbits)) : 0u) << 5) + (uint32_t)lane;
            const bool valid = a < (B >> 5) && k < (uint32_t)n;
            const float key = valid ? wt[jj] : -INFINITY;
            if (key > bt) { bt = key; bj = jj; }
        }
        old = select(bt, bj);
    }
    if (ww == 0 && lane == 0) idxs[0] = old;

    // weights below 1e-11 take the reference's double-precision key (fps.cu: sfps_key); they are rare, so the loop is
    // compiled twice and the choice is uniform over the cloud's warps
    bool slow = false;
    if (MODE == 1) {
        bool mine = false;
#pragma unroll
        for (int jj = 0; jj < RPL; ++jj) mine |= !(wt[jj] >= 1e-11f) && t[jj] != -INFINITY;
        // uniform over all warps of the cloud: the two instantiations of the loop below hold different bar.sync
        // instructions, and the warps of one cloud must meet at the same one (compute-sanitizer synccheck, r2)
        slow = W > 1 ? (__syncthreads_or(mine ? 1 : 0) != 0) : __any_sync(0xffffffffu, mine);
    }
    // in-lane arg-max in register (= tie) order, as NQ independent strict-'>' scans merged in order: a later segment
    // only replaces an earlier one when strictly larger, so the first maximum still wins
    constexpr int NQ = 4, QL = RPL / NQ;
    auto sample = [&](auto slow_c) {
        constexpr bool SLOW = decltype(slow_c)::value;
        const float x1 = sxyz[old * 3 + 0], y1 = sxyz[old * 3 + 1], z1 = sxyz[old * 3 + 2];
        float pbt[NQ];
        int pbj[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { pbt[q] = -INFINITY; pbj[q] = q * QL; }
#pragma unroll
        for (int jj = 0; jj < RPL; ++jj) {
            const float d = sqdist(x[jj], y[jj], z[jj], x1, y1, z1);
            const float tt = fminf(d, t[jj]);
            t[jj] = tt;
            float key = tt;
            if (MODE == 1) key = SLOW ? small_sfps_key(tt, wt[MODE == 1 ? jj : 0]) : __fmul_rn(tt, wt[MODE == 1 ? jj : 0]);
            if (key > pbt[jj / QL]) { pbt[jj / QL] = key; pbj[jj / QL] = jj; }
        }
#pragma unroll
        for (int q = 1; q < NQ; q += 2)
            if (pbt[q] > pbt[q - 1]) { pbt[q - 1] = pbt[q]; pbj[q - 1] = pbj[q]; }
        if (pbt[2] > pbt[0]) { pbt[0] = pbt[2]; pbj[0] = pbj[2]; }
        old = select(pbt[0], pbj[0]);
    };
    for (int it = first_it; it < m; ++it) {
        if (MODE == 1 && slow) sample(std::true_type{});
        else sample(std::false_type{});
        if (ww == 0 && lane == 0) idxs[it] = old;
    }

    // ---- the running min-distances are an in/out tensor of the op -------------------------------------------------
#pragma unroll
    for (int jj = 0; jj < RPL; ++jj) {
        const uint32_t J = (uint32_t)(ww * RPL + jj);
        const uint32_t a = J / (uint32_t)C, c = J - a * (uint32_t)C;
        const uint32_t k = c * B + ((hbits ? (__brev(a) >> (32 - hbits)) : 0u) << 5) + (uint32_t)lane;
        if (a < (B >> 5) && k < (uint32_t)n) temp_g[k] = t[jj];
    }
}

template <int MODE, int W, int RPL>
static int fs_launch(int b, int n, int m, int log2B, int C, const float *xyz, const float *w, float *temp, int *idx,
                     cudaStream_t s) {
    constexpr int CPB = 4 / W;
    const size_t smem = (size_t)CPB * n * 3 * sizeof(float);
    static unsigned long long devs = 0;   // up to 48 KB dynamic + the static exchange slots: needs the opt-in
    if (int rc = de6d_ensure_smem(fps_small_kernel<MODE, W, RPL>, 64 * 1024, devs, "fps_small smem attribute")) return rc;
    const int cinv = (65536 + C - 1) / C;
    fps_small_kernel<MODE, W, RPL><<<ceil_div(b, CPB), FS_THREADS, smem, s>>>(b, n, m, log2B, C, cinv, xyz, w, temp, idx);
    DE6D_CHECK_LAUNCH("fps_small_kernel");
    return DE6D_OK;
}

// mode 0 = D-FPS, 1 = S-FPS.  Returns -1 when the shape is outside this kernel's range (caller falls through).
int fps_small_dispatch(int mode, int b, int n, int m, int log2B, const float *xyz, const float *w, float *temp, int *idx,
                       cudaStream_t s) {
    if (n < 32 || n > 4096 || log2B < 5) return -1;
    const int B = 1 << log2B;
    const int C = (n + B - 1) / B;
    const int need = (B / 32) * C;      // slots per lane over the whole cloud
    // S-FPS keeps five values per point: 32 points per lane would spill (measured slower than the bucket kernel), so
    // it stops at 16 per lane (2048 points); D-FPS goes to 32 per lane.
#define DE6D_FS(W_, R_)                                                                                   \
    return mode == 0 ? fs_launch<0, W_, R_>(b, n, m, log2B, C, xyz, w, temp, idx, s)                      \
                     : fs_launch<1, W_, R_>(b, n, m, log2B, C, xyz, w, temp, idx, s)
    if (need <= 16) { DE6D_FS(1, 16); }
    if (need <= 32 && mode == 0) return fs_launch<0, 1, 32>(b, n, m, log2B, C, xyz, w, temp, idx, s);
    if (need <= 64) { DE6D_FS(4, 16); }
    if (need <= 128 && mode == 0) return fs_launch<0, 4, 32>(b, n, m, log2B, C, xyz, w, temp, idx, s);
#undef DE6D_FS
    return -1;
}

}  // namespace de6d
