// Fused set-abstraction scale: grouping -> shared MLP (1x1 conv + folded BatchNorm + ReLU, up to 4 layers) -> idx_cnt mask
// -> max-pool over nsample, on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators and inter-layer
// activations in TENSOR MEMORY), so that neither the grouped tensor (B, 3+C, npoint, nsample) nor any hidden activation
// ever reaches HBM.  SURVEY.md 8(f) rank 1, second half.
//
// Replaces the per-scale body of _PointnetSAModuleFSBase.forward (pointnet2/pointnet2_batch/pointnet2_modules.py:461-478):
//     idx_cnt, new_features = self.groupers[i](xyz, new_xyz, features)      # ball query + group + centre + cat
//     new_features = self.mlps[i](new_features)                             # [Conv2d(k=1, bias=False), BatchNorm2d, ReLU] x L
//     new_features *= (idx_cnt > 0)                                         # empty balls -> 0
//     pooled = F.max_pool2d(new_features, kernel_size=[1, nsample])         # (B, C_out, npoint)
// The ball query itself stays the existing kernel (its idx / idx_cnt are inputs here).
//
// Mapping.  A tile is 128 grouped points (128 / nsample consecutive queries of one cloud) = the M dimension of one
// tcgen05.mma (cta_group::1, M = 128): TMEM lane r <-> grouped point r <-> threads r and r + 128 of the 256-thread CTA (the two
// split the columns).
//   gather   thread r reads its point's feature row (point-major copy of the features, (B, N, C): one contiguous row per
//            gathered point instead of C sectors) and its centred coordinates and writes them with tcgen05.st straight into
//            TMEM as the A operand (fp32 bits; the tensor core reads the upper 19 = tf32 by truncation, as for cuDNN's TF32
//            convolutions) -- the gathered tile never exists in shared memory either;
//   layer l  one thread issues K_l / 8 tcgen05.mma (A from TMEM, B = the layer's weights resident in shared memory as a
//            SWIZZLE_128B K-major image, D in TMEM), tcgen05.commit -> mbarrier;
//   epilogue every thread reads its lane of D with tcgen05.ld, adds the folded BN bias, applies ReLU and
//            stores it back IN PLACE as the next layer's A operand; after the last layer it applies the idx_cnt mask and
//            reduces over the nsample lanes of each query (REDUX on the float bits -- the values are >= 0 after ReLU --
//            or shuffles for nsample < 32, shared-memory atomicMax across warps for nsample > 32);
//   output   32 queries x C_out are staged in shared memory and written as 128-byte runs of (B, C_out, npoint).
// Weights of all layers stay resident in shared memory for the lifetime of a persistent CTA (TMA bulk copy once).  MLPs whose
// tf32 weights exceed one SM's shared memory (131 -> 128 -> 128 -> 256: 272 KB) run as cta_group::2 PAIRS: a cluster of two
// CTAs walks its tiles in lockstep, each CTA holds half of every layer's output rows, the leader issues one M = 256 MMA per
// k-step for both tiles (UTCHMMA.2CTA, commit multicast to both CTAs' barriers).  Beyond two SMs' shared memory (or wider than
// 256 channels) the shape is refused (DE6D_ERR_INVALID) and the caller keeps the unfused composition.  TMEM columns are
// ping-ponged between two buffers (input / output of a layer).
//
// Numerics: tf32 operands (10-bit mantissa; weights rounded to nearest when packed, activations truncated by the tensor core),
// fp32 accumulation -- the arithmetic of the reference's own
// default (torch.backends.cudnn.allow_tf32 = True for Conv2d); BN is folded into the weights before the rounding.
#include "common.cuh"

namespace de6d {

constexpr int SM_MAX_LAYERS = 4;
constexpr int SM_QTC = 32;         // queries per work item (output staging: C_out x 32 floats, 128-byte rows); 16 when that
                                   // lets a second CTA fit on the SM

struct SaMlpParams {
    int b, n, m, ns, c_feat;
    int n_layers;
    int width[SM_MAX_LAYERS + 1];      // width[0] = padded input width (multiple of 8), width[l] = channels after layer l
    int kblocks[SM_MAX_LAYERS];        // 32-column blocks of layer l's weight image
    int w_off[SM_MAX_LAYERS];          // byte offset of layer l's image in the packed buffer (multiple of 1024)
    int b_off[SM_MAX_LAYERS];          // float offset of layer l's bias
    int w_bytes, bias_floats;
    int tmem_cols;                     // power of two >= 32
    int qtc;                           // queries per work item (32 or 16)
    int pair;                          // 1: two CTAs (cta_group::2) share every layer's weights, half the output rows each
    int col[SM_MAX_LAYERS + 1];        // TMEM column of the input tile and of each layer's output
    const float *xyz, *new_xyz, *feats_pm, *w_packed, *bias;
    const int *idx, *idx_cnt;
    float *out;
    int out_c_total, out_c_off;        // out is (b, out_c_total, m); this launch writes channels [out_c_off, out_c_off + width[n_layers])
    int *status;                       // optional debug word (bounded waits), may be null
};

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n
__host__ __device__ inline uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of element (r, k) of an [R x K] fp32 matrix in the SWIZZLE_128B K-major image (blocks of 32 columns)
__host__ __device__ inline uint32_t sw128_off(int R, int r, int k) {
    const int kb = k >> 5, c = (k & 31) >> 2;
    return (uint32_t)kb * (uint32_t)R * 128u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4) +
           (uint32_t)(k & 3) * 4u;
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cta_group::2 forms: one MMA covers the tiles of BOTH CTAs of a pair (M = 256); each CTA's shared memory holds half of the
// layer's output rows (the hardware reads B from both), each CTA's tensor memory receives its own 128 rows of D
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {   // arrives on the barrier at this offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__host__ __device__ inline uint32_t umma_idesc_tf32_pair(int n) {     // M = 256 over the pair
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
// bounded wait (~ seconds): a wrong descriptor must not hang the GPU; returns false on time-out
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (long long it = 0; it < 400000000ll; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
}
// (no "memory" clobber on the tensor-memory stores: they touch no C++ object, their order against tcgen05.wait::st and the
// fences is kept by `volatile`, and without the clobber the compiler may keep independent global loads in flight across them)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}

// Packs one layer's BN-folded weights (row-major [n_out x k_in], device) into the tf32-rounded SWIZZLE_128B image the
// kernel's B descriptors expect.  `xyz_last`: the reference's channel order is (dx, dy, dz, features...) (pointnet2_utils.py:417:
// cat([grouped_xyz, grouped_features])); the kernel's A tile is (features..., dx, dy, dz, 0-pad), so layer 0's columns are rotated.
__global__ void sa_mlp_pack_kernel(int n_out, int k_in, int k_pad32, int xyz_last, int halves, size_t half_stride_bytes,
                                   const float *__restrict__ w, float *__restrict__ image) {
    const int total = n_out * k_pad32;
    const int rows_per = n_out / halves;                 // halves = 2: rows [0, n/2) go to CTA 0's image, the rest to CTA 1's
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / k_pad32, k = i - r * k_pad32;
        float v = 0.f;
        if (k < k_in) {
            const int src = xyz_last ? (k < k_in - 3 ? k + 3 : k - (k_in - 3)) : k;
            v = to_tf32(w[(size_t)r * k_in + src]);
        }
        const int h = r / rows_per, rl = r - h * rows_per;
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(image) + (size_t)h * half_stride_bytes + sw128_off(rows_per, rl, k)) = v;
    }
}

// T = 256 threads: warps w and w + 4 serve the same quarter of the tensor-memory lanes (a warp may only touch lanes
// 32 * (w % 4) ...) and split the COLUMNS of every gather / epilogue between them, so 16 warps per SM are in flight at two
// CTAs per SM.  MINB = CTAs per SM the register allocation must allow (4 for small nets, 2 for nets that fill tensor memory).
constexpr int SM_T = 256;

// PAIR: a cluster of two CTAs walks its tiles in lockstep; the leader (cluster rank 0) issues one cta_group::2 MMA per
// k-step for both tiles, and every block-wide barrier of the single-CTA form becomes a cluster barrier.
template <int MINB, bool PAIR>
__global__ void __launch_bounds__(SM_T, MINB)
sa_mlp_kernel(const SaMlpParams p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_fail;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    auto tile_sync = [&]() { if (PAIR) cluster_sync_all(); else __syncthreads(); };
    const int quarter = warp & 3, half = warp >> 2;          // lane quarter of the tile, column half of the work
    const int row = quarter * 32 + lane;                     // this thread's grouped point within a tile
    unsigned char *sW = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-byte alignment
    float *sBias = reinterpret_cast<float *>(sW + p.w_bytes);
    const int c_out = p.width[p.n_layers];
    uint32_t *sOut = reinterpret_cast<uint32_t *>(sBias + ((p.bias_floats + 31) & ~31));   // [c_out][p.qtc] float bits (>= 0)

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_fail = 0;
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    for (int i = tid; i < p.bias_floats; i += SM_T) sBias[i] = p.bias[i];
    for (int i = tid; i < c_out * p.qtc; i += SM_T) sOut[i] = 0u;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile_sync();       // PAIR: the peer's barriers exist before the leader's first multicast commit can arrive on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {   // all layers' weight images (PAIR: this CTA's half of every layer): TMA bulk copies onto one transaction barrier
        const unsigned char *wsrc = reinterpret_cast<const unsigned char *>(p.w_packed) + (size_t)rank * p.w_bytes;
        mbar_expect_tx(&bar_w, (uint32_t)p.w_bytes);
        for (int off = 0; off < p.w_bytes; off += 32768) {
            const int len = min(32768, p.w_bytes - off);
            tma_bulk_g2s(sW + off, wsrc + off, (uint32_t)len, &bar_w);
        }
    }
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
    bool ok = mbar_wait_bounded(&bar_w, 0);
    uint32_t mma_phase = 0;

    const int wpb = (p.m + p.qtc - 1) / p.qtc;               // work items per cloud
    const int n_work = p.b * wpb;
    const int tiles = p.qtc * p.ns / 128;                     // tiles per work item (host guarantees divisibility)
    const int C = p.c_feat;
    const int K = C + 3;
    // this thread's share of the input columns: [cs, ce), split at a multiple of 32 (or 8 for narrow inputs)
    const int W0 = p.width[0];
    const int split = W0 >= 64 ? ((W0 / 2) & ~31) : ((W0 / 2 + 7) & ~7);
    const int cs = half ? split : 0, ce = half ? W0 : split;

    for (int w0 = PAIR ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x; w0 < n_work && (PAIR || ok); w0 += gridDim.x) {
        // PAIR: the two CTAs take work items w0 and w0 + 1 and must execute the same barriers: a CTA without a work item of
        // its own recomputes the last one and drops the result; after a time-out nothing is waited for any more
        int work = PAIR ? w0 + (int)rank : w0;
        const bool work_valid = work < n_work;
        if (!work_valid) work = n_work - 1;
        const int bi = work / wpb, q0 = (work - bi * wpb) * p.qtc;
        int k_next;
        {
            const int ql0 = row / p.ns, qq = q0 + ql0;
            k_next = qq < p.m ? __ldg(p.idx + ((size_t)bi * p.m + qq) * p.ns + (row - ql0 * p.ns)) : 0;
        }
        for (int t = 0; t < tiles; ++t) {
            // ---- gather: grouped point `row` -> TMEM lane `row`, this thread's columns of [col[0], col[0] + width[0]) ----
            const int g = t * 128 + row;
            const int ql = g / p.ns;
            const int q = q0 + ql;
            const bool valid = q < p.m;
            int k = k_next;
            {   // the next tile's index is requested now, a whole tile of work before it is needed
                const int g2 = (t + 1) * 128 + row, ql2 = g2 / p.ns, q2 = q0 + ql2;
                k_next = (t + 1 < tiles && q2 < p.m) ? __ldg(p.idx + ((size_t)bi * p.m + q2) * p.ns + (g2 - ql2 * p.ns)) : 0;
            }
            k = min(max(k, 0), p.n - 1);
            const float *frow = p.feats_pm + ((size_t)bi * p.n + k) * C;
            float rel[3] = {0.f, 0.f, 0.f};
            if (ce > C) {   // only the half that owns the coordinate columns needs them
                const float *prow = p.xyz + ((size_t)bi * p.n + k) * 3;
                const float *crow = p.new_xyz + ((size_t)bi * p.m + (valid ? q : 0)) * 3;
                rel[0] = __fsub_rn(__ldg(prow), __ldg(crow)); rel[1] = __fsub_rn(__ldg(prow + 1), __ldg(crow + 1));
                rel[2] = __fsub_rn(__ldg(prow + 2), __ldg(crow + 2));
            }
            const bool vec = (C & 3) == 0;
            // feature columns in batches of 32: eight 16-byte loads are issued before the first store, so a row costs
            // ~C/32 L2 round trips instead of C/8.  Values go to tensor memory as fp32 bits: the tensor core reads the upper
            // 19 bits (tf32 by truncation), which is what the reference's cuDNN TF32 convolutions feed it as well.
            int c0 = cs;
            if (vec) {
                for (; c0 + 32 <= min(ce, C); c0 += 32) {
                    float4 f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = __ldg(reinterpret_cast<const float4 *>(frow + c0) + j);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        uint32_t v[8];
                        v[0] = __float_as_uint(f[2 * h].x); v[1] = __float_as_uint(f[2 * h].y); v[2] = __float_as_uint(f[2 * h].z); v[3] = __float_as_uint(f[2 * h].w);
                        v[4] = __float_as_uint(f[2 * h + 1].x); v[5] = __float_as_uint(f[2 * h + 1].y); v[6] = __float_as_uint(f[2 * h + 1].z); v[7] = __float_as_uint(f[2 * h + 1].w);
                        tmem_st8(lane_addr + (uint32_t)(p.col[0] + c0 + 8 * h), v);
                    }
                }
            }
            for (; c0 < ce; c0 += 8) {
                uint32_t v[8];
                if (vec && c0 + 8 <= C) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(frow + c0));
                    const float4 b2 = __ldg(reinterpret_cast<const float4 *>(frow + c0 + 4));
                    v[0] = __float_as_uint(a.x); v[1] = __float_as_uint(a.y); v[2] = __float_as_uint(a.z); v[3] = __float_as_uint(a.w);
                    v[4] = __float_as_uint(b2.x); v[5] = __float_as_uint(b2.y); v[6] = __float_as_uint(b2.z); v[7] = __float_as_uint(b2.w);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c0 + j;
                        float x = 0.f;
                        if (c < C) x = __ldg(frow + c);
                        else if (c < K) x = rel[c - C];
                        v[j] = __float_as_uint(x);
                    }
                }
                tmem_st8(lane_addr + (uint32_t)(p.col[0] + c0), v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

            // ---- the layers ----
            for (int l = 0; l < p.n_layers; ++l) {
                const int kin = p.width[l], nout = p.width[l + 1];
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tile_sync();
                if (tid == 0 && rank == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t idesc = PAIR ? umma_idesc_tf32_pair(nout) : umma_idesc_tf32(nout);
                    const uint32_t wbase = smem_u32(sW + p.w_off[l]);
                    const uint32_t d_t = tbase + (uint32_t)p.col[l + 1], a_t = tbase + (uint32_t)p.col[l];
                    const uint32_t brows = (uint32_t)(PAIR ? nout / 2 : nout);      // rows of B held by one CTA
                    for (int ks = 0; ks < kin / 8; ++ks) {
                        const uint32_t boff = (uint32_t)(ks >> 2) * brows * 128u + (uint32_t)(ks & 3) * 32u;
                        if (PAIR) umma_ts_pair(d_t, a_t + (uint32_t)ks * 8u, umma_desc_sw128(wbase + boff), idesc, ks > 0);
                        else umma_ts(d_t, a_t + (uint32_t)ks * 8u, umma_desc_sw128(wbase + boff), idesc, ks > 0);
                    }
                    if (PAIR) umma_commit_pair(&bar_mma); else umma_commit(&bar_mma);
                }
                if (ok) ok = mbar_wait_bounded(&bar_mma, mma_phase);
                mma_phase ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (!ok) { s_fail = 1; if (!PAIR) break; }
                const float *bias = sBias + p.b_off[l];
                if (l + 1 < p.n_layers) {
                    // hidden layer: bias + ReLU in place (next layer's A operand); 16-column chunks alternate between the halves
                    for (int c0 = half * 16; c0 < nout; c0 += 32) {
                        uint32_t v[16];
                        float bb[16];
                        tmem_ld16(lane_addr + (uint32_t)(p.col[l + 1] + c0), v);
#pragma unroll
                        for (int j = 0; j < 4; ++j)    // four broadcast 16-byte loads while the tensor-memory load is in flight
                            *reinterpret_cast<float4 *>(bb + 4 * j) = *reinterpret_cast<const float4 *>(bias + c0 + 4 * j);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]) + bb[j], 0.f));
                        tmem_st16(lane_addr + (uint32_t)(p.col[l + 1] + c0), v);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                } else {
                    // last layer: bias + ReLU + mask, max over the nsample lanes of each query
                    const bool live = valid && (p.idx_cnt == nullptr || p.idx_cnt[(size_t)bi * p.m + q] > 0);
                    const int seg = p.ns < 32 ? p.ns : 32;                 // lanes of this warp that share a query
                    const bool leader = (lane & (seg - 1)) == 0;
                    for (int c0 = half * 16; c0 < nout; c0 += 32) {
                        uint32_t v[16];
                        float bb[16];
                        tmem_ld16(lane_addr + (uint32_t)(p.col[l + 1] + c0), v);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4 *>(bb + 4 * j) = *reinterpret_cast<const float4 *>(bias + c0 + 4 * j);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float y = live ? fmaxf(__uint_as_float(v[j]) + bb[j], 0.f) : 0.f;
                            uint32_t u = __float_as_uint(y);               // y >= +0: unsigned order == float order
                            if (seg == 32) {
                                u = __reduce_max_sync(0xffffffffu, u);
                            } else {
                                for (int o = seg >> 1; o; o >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, o));
                            }
                            v[j] = u;
                        }
                        if (leader && valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) atomicMax(&sOut[(c0 + j) * p.qtc + ql], v[j]);
                        }
                    }
                }
            }
            if (!ok && !PAIR) break;
        }
        // ---- write the work item's 32 queries x C_out, re-arm the staging buffer ----
        __syncthreads();
        if (ok && !s_fail && work_valid) {
            const int nq = min(p.qtc, p.m - q0);
            for (int i = tid; i < c_out * p.qtc; i += SM_T) {
                const int c = i / p.qtc, j = i - c * p.qtc;
                if (j < nq) p.out[((size_t)bi * p.out_c_total + p.out_c_off + c) * p.m + q0 + j] = __uint_as_float(sOut[i]);
            }
        }
        __syncthreads();
        for (int i = tid; i < c_out * p.qtc; i += SM_T) sOut[i] = 0u;
        __syncthreads();
    }
    if ((!ok || s_fail) && p.status && tid == 0) atomicExch(p.status, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile_sync();
    if (tid == 0) { mbar_inval(&bar_w); mbar_inval(&bar_mma); }     // nothing can arrive on them any more
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(p.tmem_cols));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(p.tmem_cols));
    }
}

struct SaMlpPlan {
    SaMlpParams p;
    size_t smem;
    int ctas_per_sm;
};

// widths: [c_in (= 3 + c_feat), c_1, ..., c_L].  Returns 0 and fills `plan`, or a negative reason code.
// pair = false: one CTA holds all weights.  pair = true: a cta_group::2 pair, each CTA holds half the output rows of every
// layer (widths multiples of 32: the 2-CTA MMA with A from tensor memory needs N % 32 == 0).
static int sa_mlp_plan_one(int n_layers, const int *widths, int ns, bool pair, SaMlpPlan &plan) {
    SaMlpParams &p = plan.p;
    if (n_layers < 1 || n_layers > SM_MAX_LAYERS) return -1;
    if (!(ns == 4 || ns == 8 || ns == 16 || ns == 32 || ns == 64 || ns == 128)) return -2;      // lanes per query must tile 128
    p.n_layers = n_layers;
    p.pair = pair ? 1 : 0;
    p.width[0] = (widths[0] + 7) & ~7;
    int woff = 0, boff = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int nout = widths[l + 1];
        if (nout < 16 || nout > 256 || (nout & (pair ? 31 : 15))) return -3;      // A-from-TMEM MMA: N % 16 == 0 (pair: 32), N <= 256
        p.width[l + 1] = nout;
        p.kblocks[l] = (p.width[l] + 31) / 32;
        p.w_off[l] = woff;
        p.b_off[l] = boff;
        woff += (pair ? nout / 2 : nout) * p.kblocks[l] * 128;       // multiple of 1024 because the row count is a multiple of 8
        boff += nout;
    }
    p.w_bytes = woff;                                                // per CTA
    p.bias_floats = boff;
    // TMEM: inputs / outputs alternate between two column buffers
    int buf[2] = {0, 0};
    for (int l = 0; l <= n_layers; ++l) buf[l & 1] = max(buf[l & 1], p.width[l]);
    const int b0 = (buf[0] + 15) & ~15;
    if (b0 + buf[1] > 512) return -4;
    int cols = 32;
    while (cols < b0 + buf[1]) cols <<= 1;
    p.tmem_cols = cols;
    for (int l = 0; l <= n_layers; ++l) p.col[l] = (l & 1) ? b0 : 0;
    const int by_tmem = 512 / cols;
    auto smem_for = [&](int qtc) { return 1024 + (size_t)p.w_bytes + (size_t)((p.bias_floats + 31) & ~31) * 4 + (size_t)p.width[n_layers] * qtc * 4; };
    auto ctas_for = [&](int qtc) { return max(1, min(min((int)((226 * 1024) / (smem_for(qtc) + 1024)), by_tmem), 4)); };
    p.qtc = SM_QTC;
    if ((16 * ns) % 128 == 0 && (ctas_for(16) > ctas_for(SM_QTC) || smem_for(SM_QTC) > 220 * 1024)) p.qtc = 16;   // smaller staging buffer: another CTA per SM / fits at all
    plan.smem = smem_for(p.qtc);
    if (plan.smem > 220 * 1024) return -5;                           // weights must stay resident in shared memory
    plan.ctas_per_sm = pair ? 1 : ctas_for(p.qtc);
    return 0;
}
static int sa_mlp_plan(int n_layers, const int *widths, int ns, SaMlpPlan &plan) {
    const int rc = sa_mlp_plan_one(n_layers, widths, ns, false, plan);
    if (rc != -5) return rc;
    return sa_mlp_plan_one(n_layers, widths, ns, true, plan) == 0 ? 0 : -5;   // too large for one SM: try a CTA pair
}

}  // namespace de6d

using namespace de6d;

// Can this MLP shape run fused?  widths = [3 + c_feat, c_1, ..., c_L] (host array).  Returns 0 (no), 1 (one CTA holds all
// weights) or 2 (cta_group::2 pairs: half of every layer's weights per SM).
extern "C" int de6d_sa_mlp_fits(int n_layers, const int *widths, int nsample) {
    SaMlpPlan plan;
    if (!widths || sa_mlp_plan(n_layers, widths, nsample, plan) != 0) return 0;
    return plan.p.pair ? 2 : 1;
}
// Floats of the packed weight buffer / of the bias buffer for de6d_sa_mlp_pack and de6d_sa_mlp_fused.
extern "C" size_t de6d_sa_mlp_packed_floats(int n_layers, const int *widths) {
    SaMlpPlan plan;
    if (!widths || sa_mlp_plan(n_layers, widths, 32, plan) != 0) return 0;
    return (size_t)plan.p.w_bytes / 4 * (plan.p.pair ? 2 : 1);
}
// weights_cat (device): the layers' BN-folded weight matrices, row-major [c_{l+1} x c_l], concatenated; column order of layer
// 0 as in the reference's grouped tensor (dx, dy, dz, features...).  packed (device): de6d_sa_mlp_packed_floats floats.
extern "C" int de6d_sa_mlp_pack(int n_layers, const int *widths, const float *weights_cat, float *packed, cudaStream_t stream) {
    SaMlpPlan plan;
    if (!widths || !weights_cat || !packed) return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_pack: null pointer");
    if (sa_mlp_plan(n_layers, widths, 32, plan) != 0) return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_pack: shape not supported by the fused kernel");
    size_t src = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int nout = widths[l + 1], kin = widths[l];
        const int kpad = plan.p.kblocks[l] * 32;
        sa_mlp_pack_kernel<<<ceil_div(nout * kpad, 256), 256, 0, stream>>>(nout, kin, kpad, l == 0 ? 1 : 0, plan.p.pair ? 2 : 1,
                                                                           (size_t)plan.p.w_bytes, weights_cat + src,
                                                                           packed + plan.p.w_off[l] / 4);
        DE6D_CHECK_LAUNCH("sa_mlp_pack_kernel");
        src += (size_t)nout * kin;
    }
    return DE6D_OK;
}

// xyz (b,n,3), new_xyz (b,m,3), feats_pm (b,n,c_feat) POINT-major (NULL when c_feat == 0), idx (b,m,nsample), idx_cnt (b,m) or
// NULL (no empty-ball mask), packed / bias from de6d_sa_mlp_pack (bias: concatenated folded BN biases), out (b, c_L, m).
// ..._slice: out is (b, out_channels, m) and this launch fills channels [out_channel_offset, out_channel_offset + c_L) of it.
// A last layer too wide for the shared memory of an SM pair is run as several launches over row blocks of its weight
// matrix (each launch recomputes the earlier layers: de6d_b200/sa_fused.py decides), e.g. 131 -> 128 -> 256 -> {128 | 128}.
extern "C" int de6d_sa_mlp_fused_slice(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                                       const float *feats_pm, const int *idx, const int *idx_cnt, int n_layers, const int *widths,
                                       const float *packed, const float *bias, float *out, int out_channels, int out_channel_offset,
                                       int *status, cudaStream_t stream);
extern "C" int de6d_sa_mlp_fused(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                                 const float *feats_pm, const int *idx, const int *idx_cnt, int n_layers, const int *widths,
                                 const float *packed, const float *bias, float *out, int *status, cudaStream_t stream) {
    if (!widths || n_layers < 1 || n_layers > SM_MAX_LAYERS) return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused: widths / n_layers");
    return de6d_sa_mlp_fused_slice(b, n, m, nsample, c_feat, xyz, new_xyz, feats_pm, idx, idx_cnt, n_layers, widths, packed, bias, out,
                                   widths[n_layers], 0, status, stream);
}
extern "C" int de6d_sa_mlp_fused_slice(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                                       const float *feats_pm, const int *idx, const int *idx_cnt, int n_layers, const int *widths,
                                       const float *packed, const float *bias, float *out, int out_channels, int out_channel_offset,
                                       int *status, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0 || nsample < 0 || c_feat < 0) return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (!xyz || !new_xyz || !idx || !widths || !packed || !bias || !out || (c_feat > 0 && !feats_pm) || n == 0)
        return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused: null pointer");
    if (widths[0] != c_feat + 3) return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused: widths[0] must be 3 + c_feat");
    SaMlpPlan plan;
    if (sa_mlp_plan(n_layers, widths, nsample, plan) != 0)
        return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused: shape not supported (layer widths multiple of 16 and <= 256, weights resident in shared memory, nsample a power of two in 4..128)");
    if ((reinterpret_cast<uintptr_t>(packed) & 15) || (c_feat % 4 == 0 && (reinterpret_cast<uintptr_t>(feats_pm) & 15)))
        return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused: packed weights / features must be 16-byte aligned");
    SaMlpParams &p = plan.p;
    p.b = b; p.n = n; p.m = m; p.ns = nsample; p.c_feat = c_feat;
    p.xyz = xyz; p.new_xyz = new_xyz; p.feats_pm = feats_pm; p.w_packed = packed; p.bias = bias; p.idx = idx; p.idx_cnt = idx_cnt;
    p.out = out; p.status = status;
    if (out_channel_offset < 0 || out_channel_offset + widths[n_layers] > out_channels)
        return de6d_set_error(DE6D_ERR_INVALID, "sa_mlp_fused_slice: channel slice outside the output");
    p.out_c_total = out_channels; p.out_c_off = out_channel_offset;
    static unsigned long long devs[3] = {0, 0, 0};
    const long long n_work = (long long)b * ceil_div(m, p.qtc);
    if (p.pair) {
        // one cluster of two CTAs per SM pair; the pair takes two work items per pass
        if (int rc = de6d_ensure_smem(sa_mlp_kernel<1, true>, 226 * 1024, devs[2], "sa_mlp smem attribute")) return rc;
        long long pairs = (n_work + 1) / 2;
        if (pairs > 74) pairs = 74;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(SM_T);
        cfg.dynamicSmemBytes = plan.smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, sa_mlp_kernel<1, true>, p);
        if (e != cudaSuccess) return de6d_set_cuda_error(e, "sa_mlp_kernel (pair) launch");
        DE6D_CHECK_LAUNCH("sa_mlp_kernel (pair)");
        return DE6D_OK;
    }
    const bool small = plan.ctas_per_sm >= 3;      // 4 CTAs of 256 threads per SM: <= 64 registers per thread
    if (int rc = small ? de6d_ensure_smem(sa_mlp_kernel<4, false>, 226 * 1024, devs[0], "sa_mlp smem attribute")      // 227 KB minus the static barriers
                       : de6d_ensure_smem(sa_mlp_kernel<2, false>, 226 * 1024, devs[1], "sa_mlp smem attribute")) return rc;
    long long grid = 148ll * plan.ctas_per_sm;
    if (grid > n_work) grid = n_work;
    if (small) sa_mlp_kernel<4, false><<<(unsigned)grid, SM_T, plan.smem, stream>>>(p);
    else sa_mlp_kernel<2, false><<<(unsigned)grid, SM_T, plan.smem, stream>>>(p);
    DE6D_CHECK_LAUNCH("sa_mlp_kernel");
    return DE6D_OK;
}
