// Farthest point sampling for sm_100a: D-FPS, S-FPS (semantics weighted) and F-FPS (distance matrix).
//
// Replaces pointnet2_batch/src/sampling_gpu.cu:101-585 of the reference (one strided CTA per cloud that
// re-reads xyz and temp from global memory every iteration and reduces through a 10-level shared-memory
// tree).  Results are bit-identical to that kernel, including its tie rule (common.cuh: fps_prio).
//
// Design (DESIGN.md section "FPS"):
//   * one CTA per cloud, the whole cloud resident on chip: coordinates SoA in shared memory (192 KB for
//     16384 points), running min-distances and tie priorities in registers;
//   * points are Morton-sorted once per launch into buckets of 32 (one warp lane per point); every bucket
//     keeps its bounding box, its largest min-distance and its cached arg-max.  A bucket whose box is
//     farther from the newly selected point than its largest min-distance cannot change and is skipped.
//     The lower bound is evaluated with the same rounded operations as the distance itself and every
//     operation involved is monotone, so skipping is exact, not approximate (see bucket_lower_bound);
//   * the per-iteration arg-max is REDUX (redux.sync) on order-preserving integer images of the floats:
//     bucket -> warp -> CTA with ONE __syncthreads per selected point (double-buffered exchange slots).
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <type_traits>

namespace de6d {

enum { FPS_D = 0, FPS_S = 1 };

// Explicit shared-window accesses with a 32-bit address computed once: generic pointers derived from the dynamic
// shared array otherwise cost an S2R SR_CgaCtaId + LEA (shared-window base) in front of every access of the loop.
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32x2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

// ---- thread-block-cluster exchange (clouds larger than one SM: one CTA per 16384-point slice) ----------------
__device__ __forceinline__ uint32_t fps_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void fps_st_async(uint32_t remote_addr, uint32_t v, uint32_t remote_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v), "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void fps_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fps_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fps_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
    // the spin loop lives inside one asm block, so the compiler places no reconvergence point behind it: make the lanes
    // meet again explicitly before anything warp-synchronous (REDUX, bar.sync) follows
    __syncwarp();
}

// S-FPS key: reference evaluates d * max(w, 1e-12) in double and rounds once to float
// (sampling_gpu.cu:465; 1e-12 is a double literal).  For w >= 1e-12f the double product of two floats is
// exact, so one float multiply gives the same rounding; only smaller weights take the double path.
__device__ __forceinline__ float sfps_key(float d, float w) {
    if (w >= 1e-11f) return __fmul_rn(d, w);
    return (float)((double)d * fmax((double)w, 1e-12));
}

// CL = true: the cloud is split over the CTAs of a thread-block cluster, CAP points each (n_in up to 8 * 16384).  Every
// CTA runs the same bucket machinery on its slice; per sample the CTAs exchange one 5-word candidate row (value,
// global priority, x, y, z) with remote stores that complete on a transaction mbarrier of the destination CTA, and
// every CTA picks the winner locally -- the selected point's coordinates arrive with the row.
//
// SPECK > 1 (D-FPS, single CTA): up to SPECK samples per barrier round, still the exact reference sequence.  Let
// c1 > c2 > ... be the points ranked by (min-distance desc, tie priority asc) at the start of a round.  c1 is the next
// sample.  Adding c1 can only lower min-distances, so if c2's own min-distance is untouched by c1 (d(c1,c2) >= temp[c2]
// with the kernel's rounded distance, temp[c2] > 0) then c2 is still the maximum afterwards, i.e. it IS the sample
// after c1; likewise c3 if untouched by c1 and c2, and so on.  The round therefore takes the top-SPECK candidates,
// accepts the longest prefix that passes those pairwise tests, and applies all accepted updates in ONE pass over
// the (pruned) buckets with ONE barrier.  Candidates come from the per-bucket cached maxima; a point hidden behind
// its bucket's maximum or behind a warp's two reported maxima could out-rank a candidate, so every bucket also
// caches its second-best value and every warp reports the largest value it did not report: a candidate is only
// accepted if it is strictly above that bound.  On FPS workloads ~3.5 of 4 candidates are accepted per round.
#ifdef DE6D_FPS_STATS   // tuning builds only (scripts/micro/fps_stats.py): rounds, accepted samples, why a round stopped
__device__ unsigned long long de6d_fps_stats_dev[8];
#endif

template <int MODE, int NW, int BPW, bool PRUNE, bool CL = false, int SPECK = 1>
__global__ void __launch_bounds__(NW * 32, 1)
fps_bucket_kernel(int n_in, int m, int log2B, int ibits, const float *__restrict__ xyz_all,
                  const float *__restrict__ w_all, float *__restrict__ temp_all, int *__restrict__ idx_all) {
    constexpr int T = NW * 32;
    constexpr int CAP = NW * BPW * 32;
    static_assert(BPW <= 32, "one owner lane per bucket");
    static_assert(!CL || MODE == FPS_D, "the cluster variant covers D-FPS");
    static_assert(SPECK == 1 || (MODE == FPS_D && !CL && PRUNE && 2 * NW <= 32), "multi-sample rounds: single-CTA pruned D-FPS");
    // tie priorities in shared memory instead of registers: the multi-sample variant, and S-FPS with 32 slots per lane
    // (min-distances + weights + packed priorities would not fit 128 registers)
    constexpr bool SCP = SPECK > 1 || (MODE == FPS_S && BPW == 32 && !CL);
    if (m <= 0) return;
    int rank = 0, S = 1, cloud = blockIdx.x;
    if (CL) {
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        rank = (int)cluster.block_rank();
        S = (int)cluster.num_blocks();
        cloud = blockIdx.x / S;
    }
    const int lo = rank * CAP;                                  // first point of this CTA's slice
    const int n = CL ? min(max(n_in - lo, 0), CAP) : n_in;      // points this CTA owns

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + CAP;
    float *sz = sy + CAP;
    uint2 *wbuf = reinterpret_cast<uint2 *>(sz + CAP);            // [2][NW]
    float *red = reinterpret_cast<float *>(wbuf + 2 * NW);        // [6][NW] prologue reductions
    int *misc = reinterpret_cast<int *>(red + 6 * NW);            // [0]=pos of index 0, [1]=non-finite flag
    uint32_t *rows = reinterpret_cast<uint32_t *>(misc + 4);      // CL: [2][8][8] candidate rows of the cluster
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(rows + 2 * 8 * 8);   // CL: [2]
    uint4 *wq = reinterpret_cast<uint4 *>(mbar + 2);              // SPECK: [2][NW] two best bucket maxima of every warp
    uint32_t *bq = reinterpret_cast<uint32_t *>(wq + 2 * NW);     // SPECK: [2][NW] largest value a warp did not report
    unsigned short *scp = reinterpret_cast<unsigned short *>(bq + 2 * NW);   // SPECK: [CAP] tie priorities (frees 16 registers)
    unsigned long long *sortbuf = reinterpret_cast<unsigned long long *>(smem_raw);  // aliases sx/sy

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float *xyz_cloud = xyz_all + (size_t)cloud * n_in * 3;
    const float *xyz = xyz_cloud + (size_t)lo * 3;
    const float *wts = MODE == FPS_S ? w_all + (size_t)cloud * n_in : nullptr;
    float *temp_g = temp_all + (size_t)cloud * n_in + lo;
    int *idxs = idx_all + (size_t)cloud * m;

    if (tid < 2) misc[tid] = 0;
    __syncthreads();

    // ---------------- prologue: order the cloud into buckets -----------------------------------------
    uint32_t kk[BPW];  // original index of the point at (bucket j*NW+w, lane); 0xffffffff for padding
    bool prune = PRUNE;
    if (PRUNE) {
        // cloud bounding box + finiteness
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        bool bad = false;
        for (int k = tid; k < n; k += T) {
            float t0 = temp_g[k];
            bad |= (t0 != t0);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float v = xyz[k * 3 + a];
                bad |= !(fabsf(v) <= 3.0e38f);
                lo[a] = fminf(lo[a], v);
                hi[a] = fmaxf(hi[a], v);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
            if (lane == 0) { red[a * NW + w] = lo[a]; red[(3 + a) * NW + w] = hi[a]; }
        }
        if (bad) misc[1] = 1;
        __syncthreads();
        float ext = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float l = INFINITY, h = -INFINITY;
            for (int i = 0; i < NW; ++i) { l = fminf(l, red[a * NW + i]); h = fmaxf(h, red[(3 + a) * NW + i]); }
            lo[a] = l;
            ext = fmaxf(ext, h - l);
        }
        prune = (misc[1] == 0);
        const float inv = ext > 0.f ? 1023.0f / ext : 0.f;
        __syncthreads();  // red[] fully consumed before the sort buffers (aliasing only sx/sy/sz, but keep phases apart)
        // Stable LSD radix sort (6-bit digits, 5 passes over the 30-bit Morton code) where its buffers fit under the
        // coordinate arrays: the same order as sorting (code << 32 | index), i.e. the same buckets, for ~1/10 of the
        // shared-memory traffic of the bitonic network below (105 passes over 16384 8-byte items took 60 % of the
        // prologue, 0.25 ms of a 2.6 ms launch).  Warp w ranks the items at positions [w * 32 * BPW, (w + 1) * 32 * BPW):
        // per 32 items one match.any gives the rank among equal digits inside the warp, hist[digit][warp] the count in
        // the warp's earlier items; an exclusive scan over (digit, warp) turns the counts into stable global offsets.
        constexpr bool USE_RADIX = (2 * CAP >= 260 * NW + 64) && CAP <= 16384;
        if constexpr (USE_RADIX) {
            uint32_t *keys = reinterpret_cast<uint32_t *>(smem_raw);                    // [CAP] Morton code of point k
            unsigned short *ia = reinterpret_cast<unsigned short *>(keys + CAP);        // [CAP] point at position p (ping)
            unsigned short *ib = ia + CAP;                                              // [CAP] (pong)
            unsigned short *loff = ib + CAP;                                            // [CAP] rank of position p inside (digit, warp)
            uint32_t *hist = reinterpret_cast<uint32_t *>(loff + CAP);                  // [64][NW]
            uint32_t *wtot = hist + 64 * NW;                                            // [NW]
            for (int k = tid; k < n; k += T) {
                uint32_t key = 0;
                if (prune) {
                    uint32_t qx = (uint32_t)fminf(fmaxf((xyz[k * 3 + 0] - lo[0]) * inv, 0.f), 1023.f);
                    uint32_t qy = (uint32_t)fminf(fmaxf((xyz[k * 3 + 1] - lo[1]) * inv, 0.f), 1023.f);
                    uint32_t qz = (uint32_t)fminf(fmaxf((xyz[k * 3 + 2] - lo[2]) * inv, 0.f), 1023.f);
                    key = part1by2(qx) | (part1by2(qy) << 1) | (part1by2(qz) << 2);
                }
                keys[k] = key;
                ia[k] = (unsigned short)k;
            }
            const uint32_t lt_mask = (1u << lane) - 1u;
            if (prune) {
                for (int sh = 0; sh < 30; sh += 6) {
                    __syncthreads();
                    hist[2 * tid] = 0u; hist[2 * tid + 1] = 0u;     // 64 * NW == 2 * T
                    __syncthreads();
#pragma unroll 4
                    for (int i = 0; i < BPW; ++i) {
                        const int pos = ((w * BPW + i) << 5) | lane;
                        const bool a = pos < n;
                        const uint32_t d = a ? ((keys[ia[a ? pos : 0]] >> sh) & 63u) : (64u + (uint32_t)lane);
                        const uint32_t mm = __match_any_sync(0xffffffffu, d);
                        const uint32_t rank = __popc(mm & lt_mask);
                        const uint32_t cnt = a ? hist[d * NW + w] : 0u;
                        __syncwarp();
                        if (a && rank == 0u) hist[d * NW + w] = cnt + (uint32_t)__popc(mm);
                        __syncwarp();
                        if (a) loff[pos] = (unsigned short)(cnt + rank);
                    }
                    __syncthreads();
                    {   // exclusive scan over hist[0 .. 64 * NW), two entries per thread
                        const uint32_t e0 = hist[2 * tid], e1 = hist[2 * tid + 1];
                        uint32_t inc = e0 + e1;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o) inc += t;
                        }
                        if (lane == 31) wtot[w] = inc;
                        __syncthreads();
                        uint32_t base = 0u;
                        for (int i = 0; i < w; ++i) base += wtot[i];
                        const uint32_t ex = base + inc - (e0 + e1);
                        hist[2 * tid] = ex; hist[2 * tid + 1] = ex + e0;
                    }
                    __syncthreads();
#pragma unroll 4
                    for (int i = 0; i < BPW; ++i) {
                        const int pos = ((w * BPW + i) << 5) | lane;
                        if (pos < n) {
                            const unsigned short k = ia[pos];
                            const uint32_t d = (keys[k] >> sh) & 63u;
                            ib[hist[d * NW + w] + loff[pos]] = k;
                        }
                    }
                    unsigned short *t = ia; ia = ib; ib = t;
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < BPW; ++j) {
                int p = ((j * NW + w) << 5) | lane;
                kk[j] = p < n ? (uint32_t)ia[p] : 0xffffffffu;
            }
        } else {
        int np = 64;
        while (np < n) np <<= 1;
        for (int k = tid; k < np; k += T) {
            unsigned long long item = ~0ull;
            if (k < n) {
                uint32_t key = 0;
                if (prune) {
                    uint32_t qx = (uint32_t)fminf(fmaxf((xyz[k * 3 + 0] - lo[0]) * inv, 0.f), 1023.f);
                    uint32_t qy = (uint32_t)fminf(fmaxf((xyz[k * 3 + 1] - lo[1]) * inv, 0.f), 1023.f);
                    uint32_t qz = (uint32_t)fminf(fmaxf((xyz[k * 3 + 2] - lo[2]) * inv, 0.f), 1023.f);
                    key = part1by2(qx) | (part1by2(qy) << 1) | (part1by2(qz) << 2);
                }
                item = ((unsigned long long)key << 32) | (uint32_t)k;
            }
            sortbuf[k] = item;
        }
        __syncthreads();
        for (unsigned kb = 2; kb <= (unsigned)np; kb <<= 1) {
            for (unsigned jb = kb >> 1; jb > 0; jb >>= 1) {
                for (unsigned i = tid; i < (unsigned)np / 2; i += T) {
                    unsigned a = ((i & ~(jb - 1)) << 1) | (i & (jb - 1));
                    unsigned b = a | jb;
                    unsigned long long va = sortbuf[a], vb = sortbuf[b];
                    bool up = (a & kb) == 0;
                    if ((va > vb) == up) { sortbuf[a] = vb; sortbuf[b] = va; }
                }
                __syncthreads();
            }
        }
#pragma unroll
        for (int j = 0; j < BPW; ++j) {
            int p = ((j * NW + w) << 5) | lane;
            kk[j] = p < n ? (uint32_t)sortbuf[p] : 0xffffffffu;
        }
        }
        __syncthreads();  // every item read before the coordinates overwrite the sort buffer
    } else {
#pragma unroll
        for (int j = 0; j < BPW; ++j) {
            int p = ((j * NW + w) << 5) | lane;
            kk[j] = p < n ? (uint32_t)p : 0xffffffffu;
        }
    }

    // ---------------- load the cloud: coordinates -> smem, min-dist / priority / weight -> registers -------
    float temp[BPW];
    float wt[MODE == FPS_S ? BPW : 1];
    uint32_t cpk[(BPW + 1) / 2];  // two 14-bit priorities per register
#pragma unroll
    for (int j = 0; j < (BPW + 1) / 2; ++j) cpk[j] = 0;
    // per-lane bucket state (lane j owns bucket j*NW+w)
    float blox = INFINITY, bhix = -INFINITY, bloy = INFINITY, bhiy = -INFINITY, bloz = INFINITY, bhiz = -INFINITY;
    float bmaxt = -INFINITY;      // largest min-distance inside the bucket (pruning bound)
    uint32_t bval = 0, bword = 0xffffffffu;  // cached arg-max of the bucket: ord(key) and (cprio<<14 | pos)
    uint32_t bval2 = 0;                      // SPECK: second-best value of the bucket (hidden behind bval)

#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int p = ((j * NW + w) << 5) | lane;
        const uint32_t k = kk[j];
        float x = 0.f, y = 0.f, z = 0.f, t0 = -INFINITY;
        uint32_t cp = 0x3fffu;
        if (k != 0xffffffffu) {
            x = xyz[k * 3 + 0]; y = xyz[k * 3 + 1]; z = xyz[k * 3 + 2];
            t0 = temp_g[k];
            cp = cprio_of(k, log2B, ibits);
            if (MODE == FPS_S) wt[j] = wts[k];
            if (k == 0) misc[0] = p;
        } else if (MODE == FPS_S) {
            wt[j] = 0.f;
        }
        sx[p] = x; sy[p] = y; sz[p] = z;
        temp[j] = t0;
        if (SCP) scp[p] = (unsigned short)cp;
        else cpk[j >> 1] |= cp << (16 * (j & 1));
        if (PRUNE) {
            const bool valid = k != 0xffffffffu;
            uint32_t a0 = __reduce_min_sync(0xffffffffu, valid ? f2ord(x) : 0xffffffffu);
            uint32_t a1 = __reduce_max_sync(0xffffffffu, valid ? f2ord(x) : 0u);
            uint32_t a2 = __reduce_min_sync(0xffffffffu, valid ? f2ord(y) : 0xffffffffu);
            uint32_t a3 = __reduce_max_sync(0xffffffffu, valid ? f2ord(y) : 0u);
            uint32_t a4 = __reduce_min_sync(0xffffffffu, valid ? f2ord(z) : 0xffffffffu);
            uint32_t a5 = __reduce_max_sync(0xffffffffu, valid ? f2ord(z) : 0u);
            uint32_t anyv = __ballot_sync(0xffffffffu, valid);
            if (lane == j && anyv) {
                blox = ord2f(a0); bhix = ord2f(a1); bloy = ord2f(a2); bhiy = ord2f(a3); bloz = ord2f(a4); bhiz = ord2f(a5);
            }
        }
    }

    // ---------------- first index -----------------------------------------------------------------------
    // D-FPS starts from point 0 (sampling_gpu.cu:121-123); S-FPS from argmax(weights) with the same
    // candidate rule (value must exceed -1) and tie order (:451-455).
    int par = 0;
    float x1, y1, z1;
    int first_it;
    __syncthreads();  // smem coordinates + misc[0] visible
    float gx0 = 0.f, gy0 = 0.f, gz0 = 0.f;   // CL: coordinates of global point 0 (first sample and "nothing found" fallback)
    if (MODE == FPS_D) {
        if (CL) {
            gx0 = xyz_cloud[0]; gy0 = xyz_cloud[1]; gz0 = xyz_cloud[2];
            x1 = gx0; y1 = gy0; z1 = gz0;
            if (rank == 0 && tid == 0) idxs[0] = 0;
        } else {
            const int p0 = misc[0];
            x1 = sx[p0]; y1 = sy[p0]; z1 = sz[p0];
            if (tid == 0) idxs[0] = 0;
        }
        first_it = 1;
    } else {
        uint32_t v = 0, wd = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < BPW; ++j) {
            const int p = ((j * NW + w) << 5) | lane;
            if (p < n) {
                float val = wt[j];
                uint32_t vv = (val == val) ? f2ord(val) : 0u;
                const uint32_t cpf = SCP ? (uint32_t)scp[p] : ((cpk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
                uint32_t ww = (cpf << 14) | (uint32_t)p;
                if (vv > v || (vv == v && ww < wd)) { v = vv; wd = ww; }
            }
        }
        warp_argmax(v, wd);
        if (lane == 0) wbuf[par * NW + w] = make_uint2(v, wd);
        __syncthreads();
        uint2 e = lane < NW ? wbuf[par * NW + lane] : make_uint2(0u, 0xffffffffu);
        v = e.x; wd = e.y;
        warp_argmax(v, wd);
        par ^= 1;
        int pos, k;
        if (v != 0u && ord2f(v) > -1.0f) { pos = wd & 0x3fff; k = (int)index_of_cprio(wd >> 14, log2B, ibits); }
        else { pos = misc[0]; k = 0; }
        x1 = sx[pos]; y1 = sy[pos]; z1 = sz[pos];
        if (tid == 0) idxs[0] = k;
        first_it = 1;
    }

    // cached bucket arg-max from the caller's initial temp (normally 1e10 everywhere)
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int p = ((j * NW + w) << 5) | lane;
        float t = temp[j];
        float val = MODE == FPS_S ? sfps_key(t, wt[j]) : t;
        uint32_t v = (p < n && val == val) ? f2ord(val) : 0u;
        const uint32_t cp0 = SCP ? (uint32_t)scp[p] : ((cpk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
        uint32_t wd = (cp0 << 14) | (uint32_t)p;
        uint32_t tm = __reduce_max_sync(0xffffffffu, (p < n) ? f2ord(t) : 0u);
        const uint32_t v_own = v, wd_own = wd;
        warp_argmax(v, wd);
        if (SPECK > 1) {
            const uint32_t sec = __reduce_max_sync(0xffffffffu, wd_own == wd ? 0u : v_own);
            if (lane == j) bval2 = sec;
        }
        if (lane == j) { bval = v; bword = wd; bmaxt = tm ? ord2f(tm) : -INFINITY; }
    }

    // ---------------- main loop: one selected point per iteration ------------------------------------------
    const uint32_t sx_s = (uint32_t)__cvta_generic_to_shared(sx), sy_s = (uint32_t)__cvta_generic_to_shared(sy),
                   sz_s = (uint32_t)__cvta_generic_to_shared(sz), wbuf_s = (uint32_t)__cvta_generic_to_shared(wbuf);
    const uint32_t pos0 = (uint32_t)misc[0];
    const uint32_t lane_off = (uint32_t)((w << 5) | lane) * 4u;
    constexpr uint32_t ORD_M1 = 0x407fffffu;   // f2ord(-1.0f)
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows), mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
    uint32_t phases = 0u;
    if (CL) {
        if (tid == 0) {
            fps_mbar_init(mbar_s, 1);
            fps_mbar_init(mbar_s + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cooperative_groups::this_cluster().sync();   // every CTA of the cluster runs and has its barriers initialised
    }
    if constexpr (SPECK > 1) {
        const uint32_t scp_s = (uint32_t)__cvta_generic_to_shared(scp);
        float qx[SPECK], qy[SPECK], qz[SPECK];   // samples selected in the previous round, their updates still pending
        int A = 1;
        qx[0] = x1; qy[0] = y1; qz[0] = z1;
#pragma unroll
        for (int i = 1; i < SPECK; ++i) { qx[i] = x1; qy[i] = y1; qz[i] = z1; }
        // one pass over the buckets a pending sample can change: min-distances, bucket maximum / second / bound
        // The pending samples always fill all SPECK slots: slots beyond the accepted prefix repeat the round's first sample (a
        // minimum over a set with a repeated element is the same minimum), so neither the bounds nor the visits need a per-slot
        // "is this slot live" branch.
        auto update_pass = [&]() {
            bool act = false;
            if (lane < BPW) {
                float lb;
                if constexpr (SPECK % 2 == 0) {      // two pending samples per packed evaluation
                    lb = INFINITY;
#pragma unroll
                    for (int i = 0; i < SPECK; i += 2) {
                        const float2 b2 = bucket_lower_bound2(blox, bhix, bloy, bhiy, bloz, bhiz, make_float2(qx[i], qx[i + 1]),
                                                              make_float2(qy[i], qy[i + 1]), make_float2(qz[i], qz[i + 1]));
                        lb = fminf(lb, fminf(b2.x, b2.y));
                    }
                } else {
                    lb = bucket_lower_bound(blox, bhix, bloy, bhiy, bloz, bhiz, qx[0], qy[0], qz[0]);
#pragma unroll
                    for (int i = 1; i < SPECK; ++i)
                        lb = fminf(lb, bucket_lower_bound(blox, bhix, bloy, bhiy, bloz, bhiz, qx[i], qy[i], qz[i]));
                }
                // lb < (largest min-distance of the bucket), compared on the order-preserving integer images: the bucket maximum is
                // cached as bval = f2ord(max) (0 = empty bucket, below every image of a lower bound >= +0), so no float copy of it is kept
                act = prune ? (f2ord(lb) < bval) : (((lane * NW + w) << 5) < n);
            }
            unsigned mask = __ballot_sync(0xffffffffu, act);
            // Visit the active buckets two at a time.  Only the access to the bucket's min-distance register needs the bucket
            // number as a compile-time constant, so the warp-uniform switch holds just `t = min(d, temp[j]); temp[j] = t`;
            // the coordinate loads, the distances to the pending samples and the bucket statistics run on runtime addresses
            // outside it, and the loads / REDUX chains of the two buckets overlap.
            auto touch = [&](int j, float d) -> float {
                float t = d;
                switch (j) {
#define DE6D_FPS_CASE(J) case J: if constexpr (J < BPW) { t = fminf(d, temp[J < BPW ? J : 0]); temp[J < BPW ? J : 0] = t; } break;
                    DE6D_FPS_CASE(0) DE6D_FPS_CASE(1) DE6D_FPS_CASE(2) DE6D_FPS_CASE(3) DE6D_FPS_CASE(4) DE6D_FPS_CASE(5)
                    DE6D_FPS_CASE(6) DE6D_FPS_CASE(7) DE6D_FPS_CASE(8) DE6D_FPS_CASE(9) DE6D_FPS_CASE(10) DE6D_FPS_CASE(11)
                    DE6D_FPS_CASE(12) DE6D_FPS_CASE(13) DE6D_FPS_CASE(14) DE6D_FPS_CASE(15) DE6D_FPS_CASE(16) DE6D_FPS_CASE(17)
                    DE6D_FPS_CASE(18) DE6D_FPS_CASE(19) DE6D_FPS_CASE(20) DE6D_FPS_CASE(21) DE6D_FPS_CASE(22) DE6D_FPS_CASE(23)
                    DE6D_FPS_CASE(24) DE6D_FPS_CASE(25) DE6D_FPS_CASE(26) DE6D_FPS_CASE(27) DE6D_FPS_CASE(28) DE6D_FPS_CASE(29)
                    DE6D_FPS_CASE(30) DE6D_FPS_CASE(31)
#undef DE6D_FPS_CASE
                    default: break;
                }
                return t;
            };
            while (mask) {
                const int j0 = __ffs(mask) - 1;
                mask &= mask - 1;
                const bool two = mask != 0u;
                const int j1 = two ? __ffs(mask) - 1 : j0;
                mask &= mask - 1;   // no-op on 0
                const uint32_t off0 = (uint32_t)j0 * (NW * 32u * 4u), off1 = (uint32_t)j1 * (NW * 32u * 4u);
                const float x0 = lds_f32(sx_s + lane_off + off0), y0 = lds_f32(sy_s + lane_off + off0), z0 = lds_f32(sz_s + lane_off + off0);
                const float x1 = lds_f32(sx_s + lane_off + off1), y1 = lds_f32(sy_s + lane_off + off1), z1 = lds_f32(sz_s + lane_off + off1);
                unsigned short cps0, cps1;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(cps0) : "r"(scp_s + (lane_off >> 1) + (off0 >> 1)));
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(cps1) : "r"(scp_s + (lane_off >> 1) + (off1 >> 1)));
                // distances to the pending samples: the running minimum over them first, then against the stored min-distance
                // (fminf keeps the non-NaN operand, so the result is the same set minimum as folding them one by one)
                // (the two buckets' points as one packed operand: six packed instructions per pending sample instead of twelve)
                const float2 xx = make_float2(x0, x1), yy = make_float2(y0, y1), zz = make_float2(z0, z1);
                float2 dd = sqdist2(xx, yy, zz, qx[0], qy[0], qz[0]);
                float d0 = dd.x, d1 = dd.y;
#pragma unroll
                for (int i = 1; i < SPECK; ++i) {
                    dd = sqdist2(xx, yy, zz, qx[i], qy[i], qz[i]);
                    d0 = fminf(dd.x, d0);
                    d1 = fminf(dd.y, d1);
                }
                const float t0 = touch(j0, d0);
                float t1 = t0;
                if (two) t1 = touch(j1, d1);
                const int p0 = ((j0 * NW + w) << 5) | lane, p1 = ((j1 * NW + w) << 5) | lane;
                // padding slots (p >= n) hold -inf from the start and min() keeps it there
                const uint32_t v_own0 = (p0 < n && t0 == t0) ? f2ord(t0) : 0u, v_own1 = (p1 < n && t1 == t1) ? f2ord(t1) : 0u;
                const uint32_t wd_own0 = ((uint32_t)cps0 << 14) | (uint32_t)p0, wd_own1 = ((uint32_t)cps1 << 14) | (uint32_t)p1;
                // bucket maximum, its owner (smallest word among the lanes at the maximum) and the second-best value.
                // The second-best is the maximum again when two lanes tie, else the best value below it: both
                // REDUX after the first depend on the maximum only, so they overlap instead of forming a chain.
                const uint32_t v0 = __reduce_max_sync(0xffffffffu, v_own0);
                const uint32_t v1 = __reduce_max_sync(0xffffffffu, v_own1);
                const bool top0 = v_own0 == v0, top1 = v_own1 == v1;
                const uint32_t wd0 = __reduce_min_sync(0xffffffffu, top0 ? wd_own0 : 0xffffffffu);
                const uint32_t wd1 = __reduce_min_sync(0xffffffffu, top1 ? wd_own1 : 0xffffffffu);
                const uint32_t below0 = __reduce_max_sync(0xffffffffu, top0 ? 0u : v_own0);
                const uint32_t below1 = __reduce_max_sync(0xffffffffu, top1 ? 0u : v_own1);
                const uint32_t sec0 = __popc(__ballot_sync(0xffffffffu, top0)) > 1 ? v0 : below0;
                const uint32_t sec1 = __popc(__ballot_sync(0xffffffffu, top1)) > 1 ? v1 : below1;
                if (lane == j0) { bval = v0; bword = wd0; bval2 = sec0; }
                if (two && lane == j1) { bval = v1; bword = wd1; bval2 = sec1; }
            }
        };
        const uint32_t wq_s = (uint32_t)__cvta_generic_to_shared(wq), bq_s = (uint32_t)__cvta_generic_to_shared(bq);
        int it = first_it;
#ifdef DE6D_FPS_STATS
        unsigned long long st_rounds = 0, st_acc = 0, st_bound = 0, st_pair = 0, st_limit = 0, st_zero = 0, st_full = 0;
#endif
        while (it < m) {
            update_pass();
            // ---- this warp's two best bucket maxima + the largest value it does not report ----
            uint32_t a1 = bval, b1 = bword;
            warp_argmax(a1, b1);
            const bool is1 = (bword == b1) && (b1 != 0xffffffffu);
            uint32_t a2 = is1 ? 0u : bval, b2 = is1 ? 0xffffffffu : bword;
            warp_argmax(a2, b2);
            const bool is2 = !is1 && (bword == b2) && (b2 != 0xffffffffu);
            const uint32_t bnd = __reduce_max_sync(0xffffffffu, max((is1 || is2) ? 0u : bval, bval2));
            if (lane == 0) {
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wq_s + (uint32_t)(par * NW + w) * 16u), "r"(a1), "r"(b1), "r"(a2), "r"(b2) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(bq_s + (uint32_t)(par * NW + w) * 4u), "r"(bnd) : "memory");
            }
            __syncthreads();
            // ---- global top-SPECK of the 2 * NW reported maxima, and the global bound ----
            uint32_t ev = 0u, ew = 0xffffffffu, eb = 0u;
            if (lane < 2 * NW) {
                const uint2 e = lds_u32x2(wq_s + (uint32_t)(par * NW + (lane >> 1)) * 16u + (uint32_t)(lane & 1) * 8u);
                ev = e.x; ew = e.y;
            }
            if (lane < NW) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(eb) : "r"(bq_s + (uint32_t)(par * NW + lane) * 4u));
            par ^= 1;
            const uint32_t bound = __reduce_max_sync(0xffffffffu, eb);
            uint32_t cv[SPECK], cw[SPECK];
#pragma unroll
            for (int j = 0; j < SPECK; ++j) {
                uint32_t a = ev, b = ew;
                warp_argmax(a, b);
                cv[j] = a; cw[j] = b;
                if (ew == b) { ev = 0u; ew = 0xffffffffu; }
            }
            // ---- accept the longest prefix that is provably the reference sequence ----
            // Branch-free: the coordinates of all SPECK candidates are fetched at once, every pairwise test is evaluated, and the
            // prefix length falls out of a chain of ANDs.  (A candidate is only ever tested against the candidates ranked before
            // it, which are exactly the accepted ones when its turn comes, so testing against c_i instead of "accepted q_i" is
            // the same predicate.  The sequential form -- load, test, branch per candidate -- was 11 % of the kernel's stall
            // samples.)
            const int limit = min(SPECK, m - it);
            if (cv[0] > ORD_M1) {
                float cx[SPECK], cy[SPECK], cz[SPECK];
#pragma unroll
                for (int j = 0; j < SPECK; ++j) {
                    const uint32_t pj = cw[j] == 0xffffffffu ? 0u : (cw[j] & 0x3fffu);      // an absent candidate (value 0, word ~0) reads slot 0 and is never accepted
                    cx[j] = lds_f32(sx_s + pj * 4u); cy[j] = lds_f32(sy_s + pj * 4u); cz[j] = lds_f32(sz_s + pj * 4u);
                }
                bool okj[SPECK];
                okj[0] = true;
                // (measured and dropped: the six pairwise distances of SPECK = 4 as three packed sqdist2 evaluations -- 1 % slower)
#pragma unroll
                for (int j = 1; j < SPECK; ++j) {
                    const float tj = ord2f(cv[j]);
                    bool ok = j < limit && cv[j] > bound && tj > 0.f;
#pragma unroll
                    for (int i = 0; i < j; ++i) ok = ok && !(sqdist(cx[j], cy[j], cz[j], cx[i], cy[i], cz[i]) < tj);
                    okj[j] = ok;
                }
                A = 1;
                bool go = true;
#pragma unroll
                for (int j = 1; j < SPECK; ++j) {
#ifdef DE6D_FPS_STATS
                    if (go) {
                        if (!(j < limit)) ++st_limit;
                        else if (!(cv[j] > bound)) ++st_bound;
                        else if (!(ord2f(cv[j]) > 0.f)) ++st_zero;
                        else if (!okj[j]) ++st_pair;
                    }
#endif
                    go = go && okj[j];
                    if (go) A = j + 1;
                }
#pragma unroll
                for (int j = 0; j < SPECK; ++j) {      // slots beyond the accepted prefix repeat sample 0
                    const bool live = j < A;
                    qx[j] = live ? cx[j] : cx[0]; qy[j] = live ? cy[j] : cy[0]; qz[j] = live ? cz[j] : cz[0];
                }
                // warp j records sample j: one warp writing all of them was SPECK index decodes behind every other warp at the
                // next barrier
                static_assert(SPECK <= NW, "one warp per recorded sample");
                if (lane == 0 && w < A) {
                    uint32_t cwj = cw[0];
#pragma unroll
                    for (int j = 1; j < SPECK; ++j)
                        if (w == j) cwj = cw[j];
                    idxs[it + w] = (int)index_of_cprio(cwj >> 14, log2B, ibits);
                }
            } else {   // nothing exceeds -1: the reference selects index 0
                qx[0] = lds_f32(sx_s + pos0 * 4u); qy[0] = lds_f32(sy_s + pos0 * 4u); qz[0] = lds_f32(sz_s + pos0 * 4u);
#pragma unroll
                for (int j = 1; j < SPECK; ++j) { qx[j] = qx[0]; qy[j] = qy[0]; qz[j] = qz[0]; }
                A = 1;
                if (tid == 0) idxs[it] = 0;
            }
            it += A;
#ifdef DE6D_FPS_STATS
            ++st_rounds; st_acc += A; if (A == SPECK) ++st_full;
#endif
        }
#ifdef DE6D_FPS_STATS
        if (tid == 0) {
            atomicAdd(&de6d_fps_stats_dev[0], st_rounds); atomicAdd(&de6d_fps_stats_dev[1], st_acc); atomicAdd(&de6d_fps_stats_dev[2], st_bound);
            atomicAdd(&de6d_fps_stats_dev[3], st_pair); atomicAdd(&de6d_fps_stats_dev[4], st_limit); atomicAdd(&de6d_fps_stats_dev[5], st_zero);
            atomicAdd(&de6d_fps_stats_dev[6], st_full);
        }
#endif
        // the reference applies every sample's update except the last one's: catch up on the final round
        if (A > 1) {
#pragma unroll
            for (int j = 1; j < SPECK; ++j)
                if (j >= A - 1) { qx[j] = qx[0]; qy[j] = qy[0]; qz[j] = qz[0]; }
            update_pass();
        }
    } else {
    for (int it = first_it; it < m; ++it) {
            bool act = false;
            if (lane < BPW) {
                if (prune) {
                    float lb = bucket_lower_bound(blox, bhix, bloy, bhiy, bloz, bhiz, x1, y1, z1);
                    // D-FPS: the bucket maximum is bval = f2ord(max) itself, compare the integer images (as in the multi-sample path)
                    act = MODE == FPS_D ? (f2ord(lb) < bval) : (lb < bmaxt);
                } else {
                    act = ((lane * NW + w) << 5) < n;
                }
            }
            unsigned mask = __ballot_sync(0xffffffffu, act);
            // Visit only the active buckets.  The min-distances live in registers, so the bucket number must be a
            // compile-time constant inside the body: a warp-uniform switch (one indirect branch per active bucket)
            // instead of BPW predicated copies of the body that every iteration would have to walk through.
            auto visit = [&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if constexpr (j < BPW) {
                    const int p = ((j * NW + w) << 5) | lane;
                    constexpr uint32_t off = (uint32_t)j * NW * 32u * 4u;
                    float d = sqdist(lds_f32(sx_s + lane_off + off), lds_f32(sy_s + lane_off + off), lds_f32(sz_s + lane_off + off), x1, y1, z1);
                    float t = fminf(d, temp[j]);
                    if (p >= n) t = -INFINITY;
                    temp[j] = t;
                    float val = MODE == FPS_S ? sfps_key(t, wt[MODE == FPS_S ? j : 0]) : t;
                    uint32_t v = (p < n && val == val) ? f2ord(val) : 0u;
                    const uint32_t cpv = SCP ? (uint32_t)scp[p] : ((cpk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
                    uint32_t wd = (cpv << 14) | (uint32_t)p;
                    uint32_t tm;
                    if (MODE == FPS_S) tm = __reduce_max_sync(0xffffffffu, (p < n) ? f2ord(t) : 0u);
                    warp_argmax(v, wd);
                    if (MODE == FPS_D) tm = v;
                    if (lane == j) {
                        bval = v; bword = wd;
                        if (MODE == FPS_S) bmaxt = tm ? ord2f(tm) : -INFINITY;
                    }
                }
            };
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                switch (j) {
    #define DE6D_FPS_CASE(J) case J: visit(std::integral_constant<int, J>{}); break;
                    DE6D_FPS_CASE(0) DE6D_FPS_CASE(1) DE6D_FPS_CASE(2) DE6D_FPS_CASE(3) DE6D_FPS_CASE(4) DE6D_FPS_CASE(5)
                    DE6D_FPS_CASE(6) DE6D_FPS_CASE(7) DE6D_FPS_CASE(8) DE6D_FPS_CASE(9) DE6D_FPS_CASE(10) DE6D_FPS_CASE(11)
                    DE6D_FPS_CASE(12) DE6D_FPS_CASE(13) DE6D_FPS_CASE(14) DE6D_FPS_CASE(15) DE6D_FPS_CASE(16) DE6D_FPS_CASE(17)
                    DE6D_FPS_CASE(18) DE6D_FPS_CASE(19) DE6D_FPS_CASE(20) DE6D_FPS_CASE(21) DE6D_FPS_CASE(22) DE6D_FPS_CASE(23)
                    DE6D_FPS_CASE(24) DE6D_FPS_CASE(25) DE6D_FPS_CASE(26) DE6D_FPS_CASE(27) DE6D_FPS_CASE(28) DE6D_FPS_CASE(29)
                    DE6D_FPS_CASE(30) DE6D_FPS_CASE(31)
    #undef DE6D_FPS_CASE
                    default: break;
                }
            }
            uint32_t v = bval, wd = bword;
            warp_argmax(v, wd);
            if (lane == 0) sts_u32x2(wbuf_s + (uint32_t)(par * NW + w) * 8u, v, wd);
            __syncthreads();
            uint2 e = make_uint2(0u, 0xffffffffu);
            if (lane < NW) e = lds_u32x2(wbuf_s + (uint32_t)(par * NW + lane) * 8u);
            par ^= 1;
            v = e.x; wd = e.y;
            warp_argmax(v, wd);
            // reference candidate rule: a value must exceed -1 to be selected (best starts at -1, index 0);
            // ORD_M1 = f2ord(-1.0f), and NaN keys were mapped to 0
            const bool found = v > ORD_M1;
            const uint32_t pos = found ? (wd & 0x3fffu) : pos0;
            x1 = lds_f32(sx_s + pos * 4u); y1 = lds_f32(sy_s + pos * 4u); z1 = lds_f32(sz_s + pos * 4u);
            if (!CL) {
                if (tid == 0) idxs[it] = found ? (int)index_of_cprio(wd >> 14, log2B, ibits) : 0;
            } else {
                // this CTA's candidate -> every CTA of the cluster; global priority = (bit-reversed slot, k / B) with the
                // slice offset added to the k / B field (the slice start is a multiple of B, so the slot is unchanged)
                const uint32_t cp = wd >> 14;
                const uint32_t gprio = found ? (((cp >> ibits) << 22) | ((cp & ((1u << ibits) - 1u)) + (uint32_t)(lo >> log2B))) : 0xffffffffu;
                const uint32_t gv = found ? v : 0u;
                const int rpar = (it - first_it) & 1;
                if (tid == 0) fps_mbar_expect_tx(mbar_s + 8u * rpar, (uint32_t)S * 20u);
                if (w == 0 && lane < S) {
                    const uint32_t row = fps_mapa(rows_s + (uint32_t)((rpar * 8 + rank) * 8) * 4u, (uint32_t)lane);
                    const uint32_t rbar = fps_mapa(mbar_s + 8u * rpar, (uint32_t)lane);
                    fps_st_async(row, gv, rbar);
                    fps_st_async(row + 4u, gprio, rbar);
                    fps_st_async(row + 8u, __float_as_uint(x1), rbar);
                    fps_st_async(row + 12u, __float_as_uint(y1), rbar);
                    fps_st_async(row + 16u, __float_as_uint(z1), rbar);
                }
                fps_mbar_wait(mbar_s + 8u * rpar, (phases >> rpar) & 1u);
                phases ^= 1u << rpar;
                uint32_t rv = 0u, rp = 0xffffffffu;
                float rx = 0.f, ry = 0.f, rz = 0.f;
                if (lane < S) {
                    const uint32_t *r = rows + (rpar * 8 + lane) * 8;
                    rv = r[0]; rp = r[1]; rx = __uint_as_float(r[2]); ry = __uint_as_float(r[3]); rz = __uint_as_float(r[4]);
                }
                uint32_t bv2 = rv, bp2 = rp;
                warp_argmax(bv2, bp2);
                if (bv2 > ORD_M1) {
                    const int src = __ffs(__ballot_sync(0xffffffffu, rv == bv2 && rp == bp2)) - 1;
                    x1 = __shfl_sync(0xffffffffu, rx, src); y1 = __shfl_sync(0xffffffffu, ry, src); z1 = __shfl_sync(0xffffffffu, rz, src);
                    if (rank == 0 && tid == 0) idxs[it] = (int)fps_prio_to_index(bp2, (uint32_t)log2B);
                } else {   // the reference falls back to index 0 when no value exceeds -1
                    x1 = gx0; y1 = gy0; z1 = gz0;
                    if (rank == 0 && tid == 0) idxs[it] = 0;
                }
            }
        }
    
    }

    // ---------------- write the running min-distances back (temp is an in/out tensor of the op) ---------
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int p = ((j * NW + w) << 5) | lane;
        if (p < n) {
            uint32_t cp = SCP ? (uint32_t)scp[p] : ((cpk[j >> 1] >> (16 * (j & 1))) & 0xffffu);
            temp_g[index_of_cprio(cp, log2B, ibits)] = temp[j];
        }
    }
    if (CL) {
        cooperative_groups::this_cluster().sync();   // no CTA exits while a peer may still address its shared memory
        if (tid == 0) {      // invalidate the barrier objects: the next CTA on this SM initialises new ones at the same addresses
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar_s) : "memory");
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar_s + 8u) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Generic any-N kernels (cloud does not fit on one SM): same arithmetic and tie rule, coordinates and
// min-distances stay in global memory (L2 resident).  Also the F-FPS kernel, whose per-iteration input is
// one matrix row (sampling_gpu.cu:268-373) -- no geometry, so no pruning; HBM/L2 latency bound.
// ------------------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ uint32_t block_argmax_index(uint32_t v, uint32_t prio, uint2 *wbuf, int &par, uint32_t log2B) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = T / 32;
    warp_argmax(v, prio);
    if (lane == 0) wbuf[par * NW + w] = make_uint2(v, prio);
    __syncthreads();
    uint2 e = lane < NW ? wbuf[par * NW + lane] : make_uint2(0u, 0xffffffffu);
    par ^= 1;
    v = e.x; prio = e.y;
    warp_argmax(v, prio);
    if (v != 0u && ord2f(v) > -1.0f) return fps_prio_to_index(prio, log2B);
    return 0u;
}

template <int MODE, int T>
__global__ void __launch_bounds__(T, 1)
fps_global_kernel(int n, int m, int log2B, const float *__restrict__ xyz_all, const float *__restrict__ w_all,
                  float *__restrict__ temp_all, int *__restrict__ idx_all) {
    if (m <= 0) return;
    __shared__ uint2 wbuf[2 * (T / 32)];
    const int tid = threadIdx.x;
    const float *xyz = xyz_all + (size_t)blockIdx.x * n * 3;
    const float *wts = MODE == FPS_S ? w_all + (size_t)blockIdx.x * n : nullptr;
    float *temp = temp_all + (size_t)blockIdx.x * n;
    int *idxs = idx_all + (size_t)blockIdx.x * m;
    int par = 0, old = 0, first_it = 1;
    if (MODE == FPS_S) {
        uint32_t bv = 0, bp = 0xffffffffu;
        for (int k = tid; k < n; k += T) {
            float val = wts[k];
            uint32_t v = (val == val) ? f2ord(val) : 0u, pr = fps_prio(k, log2B);
            if (v > bv || (v == bv && pr < bp)) { bv = v; bp = pr; }
        }
        old = (int)block_argmax_index<T>(bv, bp, wbuf, par, log2B);
    }
    if (tid == 0) idxs[0] = old;
    for (int it = first_it; it < m; ++it) {
        const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
        uint32_t bv = 0, bp = 0xffffffffu;
        for (int k = tid; k < n; k += T) {
            float d = sqdist(xyz[k * 3 + 0], xyz[k * 3 + 1], xyz[k * 3 + 2], x1, y1, z1);
            float t = fminf(d, temp[k]);
            temp[k] = t;
            float val = MODE == FPS_S ? sfps_key(t, wts[k]) : t;
            uint32_t v = (val == val) ? f2ord(val) : 0u, pr = fps_prio(k, log2B);
            if (v > bv || (v == bv && pr < bp)) { bv = v; bp = pr; }
        }
        old = (int)block_argmax_index<T>(bv, bp, wbuf, par, log2B);
        if (tid == 0) idxs[it] = old;
    }
}

// F-FPS.  Thread t owns columns 4*(t + c*T) .. +3 for chunk c < NCH: the running min-distances of those
// columns stay in registers when REG (n <= 4*T*NCH), one 128-bit load per chunk fetches the matrix row.
template <int T, int NCH, bool REG, bool VEC>
__global__ void __launch_bounds__(T, 1)
fps_matrix_kernel(int n, int m, int log2B, const float *__restrict__ mat_all, float *__restrict__ temp_all,
                  int *__restrict__ idx_all) {
    if (m <= 0) return;
    __shared__ uint2 wbuf[2 * (T / 32)];
    const int tid = threadIdx.x;
    const float *mat = mat_all + (size_t)blockIdx.x * n * n;
    float *temp_g = temp_all + (size_t)blockIdx.x * n;
    int *idxs = idx_all + (size_t)blockIdx.x * m;
    float tr[REG ? NCH * 4 : 1];
    if (REG) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = 4 * (tid + c * T) + q;
                tr[c * 4 + q] = k < n ? temp_g[k] : -INFINITY;
            }
    }
    int par = 0, old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int it = 1; it < m; ++it) {
        const float *row = mat + (size_t)old * n;
        uint32_t bv = 0, bp = 0xffffffffu;
        if (REG) {
            float dv[NCH * 4];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                int k0 = 4 * (tid + c * T);
                if (VEC) {
                    float4 r = k0 < n ? __ldg(reinterpret_cast<const float4 *>(row + k0)) : make_float4(0, 0, 0, 0);
                    dv[c * 4 + 0] = r.x; dv[c * 4 + 1] = r.y; dv[c * 4 + 2] = r.z; dv[c * 4 + 3] = r.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) dv[c * 4 + q] = (k0 + q) < n ? __ldg(row + k0 + q) : 0.f;
                }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int k = 4 * (tid + c * T) + q;
                    if (k < n) {
                        float t = fminf(dv[c * 4 + q], tr[c * 4 + q]);
                        tr[c * 4 + q] = t;
                        uint32_t v = (t == t) ? f2ord(t) : 0u, pr = fps_prio(k, log2B);
                        if (v > bv || (v == bv && pr < bp)) { bv = v; bp = pr; }
                    }
                }
        } else {
            for (int k = tid; k < n; k += T) {
                float t = fminf(__ldg(row + k), temp_g[k]);
                temp_g[k] = t;
                uint32_t v = (t == t) ? f2ord(t) : 0u, pr = fps_prio(k, log2B);
                if (v > bv || (v == bv && pr < bp)) { bv = v; bp = pr; }
            }
        }
        old = (int)block_argmax_index<T>(bv, bp, wbuf, par, log2B);
        if (tid == 0) idxs[it] = old;
    }
    if (REG) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = 4 * (tid + c * T) + q;
                if (k < n) temp_g[k] = tr[c * 4 + q];
            }
    }
}

// ---------------------------------- host side ----------------------------------------------------------
static int ref_log2_block(int n) {  // opt_n_threads (cuda_utils.h:10-14), same double arithmetic
    int p = (int)(log((double)n) / log(2.0));
    int v = 1 << p;
    if (v > 1024) { v = 1024; p = 10; }
    if (v < 1) { v = 1; p = 0; }
    return p;
}

template <int MODE, int NW, int BPW, bool PRUNE, int SPECK = 1>
static int launch_bucket(int b, int n, int m, int log2B, int ibits, const float *xyz, const float *w, float *temp,
                         int *idx, cudaStream_t s) {
    constexpr int CAP = NW * BPW * 32;
    size_t smem = (size_t)CAP * 12 + 2 * NW * sizeof(uint2) + 6 * NW * sizeof(float) + 16 + 2 * 8 * 8 * 4 + 16 +
                  2 * NW * 16 + 2 * NW * 4 + ((SPECK > 1 || (MODE == FPS_S && BPW == 32)) ? (size_t)CAP * 2 : 0) + 16;
    static unsigned long long devs = 0;
    if (int rc = de6d_ensure_smem(fps_bucket_kernel<MODE, NW, BPW, PRUNE, false, SPECK>, (int)smem, devs, "fps smem attribute")) return rc;
    fps_bucket_kernel<MODE, NW, BPW, PRUNE, false, SPECK><<<b, NW * 32, smem, s>>>(n, m, log2B, ibits, xyz, w, temp, idx);
    DE6D_CHECK_LAUNCH("fps_bucket_kernel");
    return DE6D_OK;
}

// D-FPS of clouds larger than one SM's shared memory: a cluster of ceil(n / 16384) <= 8 CTAs per cloud.
static int launch_bucket_cluster(int b, int n, int m, const float *xyz, float *temp, int *idx, cudaStream_t s) {
    constexpr int NW = 16, BPW = 32, CAP = NW * BPW * 32;
    const int S = (n + CAP - 1) / CAP;
    size_t smem = (size_t)CAP * 12 + 2 * NW * sizeof(uint2) + 6 * NW * sizeof(float) + 16 + 2 * 8 * 8 * 4 + 16 + 2 * NW * 20 + 16;
    auto kern = fps_bucket_kernel<FPS_D, NW, BPW, true, true>;
    static unsigned long long devs = 0;
    if (int rc = de6d_ensure_smem(kern, (int)smem, devs, "fps cluster smem attribute")) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(S * b));
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const float *w = nullptr;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, n, m, 10, 4, xyz, w, temp, idx);
    if (e != cudaSuccess) return de6d_set_cuda_error(e, "fps_bucket_kernel (cluster)");
    DE6D_CHECK_LAUNCH("fps_bucket_kernel (cluster)");
    return DE6D_OK;
}

// fps_small.cu: register-resident kernel for clouds of 32..4096 points; -1 = shape not covered
int fps_small_dispatch(int mode, int b, int n, int m, int log2B, const float *xyz, const float *w, float *temp, int *idx,
                       cudaStream_t s);

template <int MODE>
static int fps_dispatch(int b, int n, int m, const float *xyz, const float *w, float *temp, int *idx, int impl,
                        cudaStream_t s) {
    if (b < 0 || n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "fps: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (n == 0) return de6d_set_error(DE6D_ERR_INVALID, "fps: empty cloud with npoint > 0");
    if (!xyz || !temp || !idx || (MODE == FPS_S && !w)) return de6d_set_error(DE6D_ERR_INVALID, "fps: null pointer");
    const int log2B = ref_log2_block(n);
    int ibits = 0;
    while (((n - 1) >> log2B) >> ibits) ++ibits;
    const bool prune = impl != 1;
    if (impl == 0 || impl == 6) {   // small clouds: every point every sample out of registers beats the bucket machinery
        const int rc = fps_small_dispatch(MODE == FPS_S ? 1 : 0, b, n, m, log2B, xyz, w, temp, idx, s);
        if (rc != -1) return rc;
    }
    if (n <= 16384 && impl == 3 && log2B + ibits <= 14) {   // experimental: twice the warps, half the buckets per warp
        if (n <= 1024) return launch_bucket<MODE, 16, 2, true>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        if (n <= 4096) return launch_bucket<MODE, 32, 4, true>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        return launch_bucket<MODE, 32, 16, true>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
    }
    if constexpr (MODE == FPS_D) {
        // tuning: up to 6 / 8 samples per round at 16384 points (impl 7 / 8)
        if (n > 4096 && n <= 16384 && log2B + ibits <= 14 && impl == 7) return launch_bucket<MODE, 16, 32, true, 6>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        if (n > 4096 && n <= 16384 && log2B + ibits <= 14 && impl == 8) return launch_bucket<MODE, 16, 32, true, 8>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        // default D-FPS of large clouds: pruned buckets + up to 4 samples per barrier round (measured 4-10 % faster at
        // 16384 points: 3.8 samples per round, but the per-sample instruction work, not the barrier count, bounds the
        // kernel; at <= 4096 points the one-sample rounds are as fast).  impl 5 forces it at any size (tests).
        if (n <= 16384 && ((impl == 0 && n > 4096) || impl == 5) && log2B + ibits <= 14) {
            if (n <= 1024) return launch_bucket<MODE, 8, 4, true, 4>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
            if (n <= 4096) return launch_bucket<MODE, 16, 8, true, 4>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
            return launch_bucket<MODE, 16, 32, true, 4>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        }
    }
    if (n <= 16384 && impl != 2 && log2B + ibits <= 14) {   // impl 4 (and S-FPS): one sample per round
        if (n <= 1024)
            return prune ? launch_bucket<MODE, 8, 4, true>(b, n, m, log2B, ibits, xyz, w, temp, idx, s)
                         : launch_bucket<MODE, 8, 4, false>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        if (n <= 4096)
            return prune ? launch_bucket<MODE, 16, 8, true>(b, n, m, log2B, ibits, xyz, w, temp, idx, s)
                         : launch_bucket<MODE, 16, 8, false>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
        return prune ? launch_bucket<MODE, 16, 32, true>(b, n, m, log2B, ibits, xyz, w, temp, idx, s)
                     : launch_bucket<MODE, 16, 32, false>(b, n, m, log2B, ibits, xyz, w, temp, idx, s);
    }
    if (MODE == FPS_D && n > 16384 && n <= 8 * 16384 && impl != 2) return launch_bucket_cluster(b, n, m, xyz, temp, idx, s);
    fps_global_kernel<MODE, 1024><<<b, 1024, 0, s>>>(n, m, log2B, xyz, w, temp, idx);
    DE6D_CHECK_LAUNCH("fps_global_kernel");
    return DE6D_OK;
}

}  // namespace de6d

using namespace de6d;

// impl: 0 = default (bucket-pruned on-chip kernel when the cloud fits, else generic), 1 = on-chip kernel without
// pruning (every bucket visited every iteration), 2 = generic global-memory kernel.  All give identical results.
extern "C" int de6d_furthest_point_sampling_impl(int b, int n, int m, const float *xyz, float *temp, int *idx, int impl,
                                                 cudaStream_t stream) {
    return fps_dispatch<FPS_D>(b, n, m, xyz, nullptr, temp, idx, impl, stream);
}
extern "C" int de6d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                            cudaStream_t stream) {
    return fps_dispatch<FPS_D>(b, n, m, xyz, nullptr, temp, idx, 0, stream);
}
extern "C" int de6d_furthest_point_sampling_weights_impl(int b, int n, int m, const float *xyz, const float *weights,
                                                         float *temp, int *idx, int impl, cudaStream_t stream) {
    return fps_dispatch<FPS_S>(b, n, m, xyz, weights, temp, idx, impl, stream);
}
extern "C" int de6d_furthest_point_sampling_weights(int b, int n, int m, const float *xyz, const float *weights,
                                                    float *temp, int *idx, cudaStream_t stream) {
    return fps_dispatch<FPS_S>(b, n, m, xyz, weights, temp, idx, 0, stream);
}

extern "C" int de6d_furthest_point_sampling_matrix(int b, int n, int m, const float *matrix, float *temp, int *idx,
                                                   cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "fps_matrix: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (n == 0) return de6d_set_error(DE6D_ERR_INVALID, "fps_matrix: empty cloud with npoint > 0");
    if (!matrix || !temp || !idx) return de6d_set_error(DE6D_ERR_INVALID, "fps_matrix: null pointer");
    const int log2B = ref_log2_block(n);
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(matrix) & 15) == 0);
    constexpr int T = 1024;
    if (n <= 4 * T) {
        if (vec) fps_matrix_kernel<T, 1, true, true><<<b, T, 0, stream>>>(n, m, log2B, matrix, temp, idx);
        else fps_matrix_kernel<T, 1, true, false><<<b, T, 0, stream>>>(n, m, log2B, matrix, temp, idx);
    } else if (n <= 16 * T) {
        if (vec) fps_matrix_kernel<T, 4, true, true><<<b, T, 0, stream>>>(n, m, log2B, matrix, temp, idx);
        else fps_matrix_kernel<T, 4, true, false><<<b, T, 0, stream>>>(n, m, log2B, matrix, temp, idx);
    } else {
        fps_matrix_kernel<T, 1, false, false><<<b, T, 0, stream>>>(n, m, log2B, matrix, temp, idx);
    }
    DE6D_CHECK_LAUNCH("fps_matrix_kernel");
    return DE6D_OK;
}

#ifdef DE6D_FPS_STATS
extern "C" int de6d_fps_stats_read(unsigned long long *out8, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out8, de6d::de6d_fps_stats_dev, 8 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(de6d::de6d_fps_stats_dev, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif
