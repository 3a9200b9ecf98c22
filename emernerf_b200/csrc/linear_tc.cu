// Dense layers of the MLP heads on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// The reference runs every head as fp32 nn.Linear, so a single-pass TF32 MMA
// (10-bit mantissa) is not accurate enough for the 1e-4 parity bar.  Each fp32 operand is split
// into tf32 hi + tf32 lo (x = hi + lo exactly to ~2^-21) and a product is accumulated in TMEM as
//     A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          ("3xTF32", error ~1e-6 relative)
// by three tcgen05.mma.kind::tf32 instructions per 8-wide k step.
//
// Shared-memory operand layout (both operands, no swizzle, "chunk-major"): the tile is cut into
// panels of 4 consecutive k (16 bytes); inside a panel row r sits at r*16 bytes:
//     offset(r, k) = (k/4)*PANEL + r*16 + (k%4)*4
// which is the canonical K-major SWIZZLE_NONE UMMA layout with 8x16B core matrices,
// SBO = 128 B (next 8 rows) and LBO = PANEL (next 4 k).  A panel of the A tile is 128 rows =
// 2048 B, padded to 2064 B so that the 8 threads that fill one row's 128 B hit 8 distinct
// bank groups (conflict-free 16-byte stores).
//
// One CTA = 128 threads = 128 rows = 128 TMEM lanes.  Thread t owns row t in the loader, and lane
// t of the accumulator in the epilogue (tcgen05.ld 32x32b: warp w reads lanes 32w..32w+31).
// The A tile streams through a 2-stage ring of 32-wide k chunks (mbarrier, released by
// tcgen05.commit); the B operand (the layer's weights, hi and lo) is resident for the CTA's life.
//
//   forward        Y  = act(X W^T + b)       A = X[128 x k],   B = W   [n_out x k]
//   backward-data  dX = (dY * act'(Y)) W     A = dZ[128 x n_out], B = W^T [k x n_out]
//
// Roofline: these per-layer kernels are HBM-bound (read X, write Y: (k + n_out)*4 B per row); the
// tensor pipe needs 3 * ceil(k/8) MMAs of 128 x n_pad per tile.
#include "common.cuh"

namespace emer {
namespace tc {

constexpr int ROWS = 128;
constexpr int CHUNK = 32;                    // k per ring stage
constexpr int PANELS_PER_CHUNK = CHUNK / 4;  // 8
constexpr int A_PANEL = ROWS * 16 + 16;      // 2064 B (padded LBO)
constexpr int A_STAGE = PANELS_PER_CHUNK * A_PANEL;   // one of hi / lo
constexpr int STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_async_proxy() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4   [16,30) LBO>>4   [32,46) SBO>>4   [46,48) version=1   [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (bit4), a=b=tf32 (2<<7, 2<<10),
// K-major both, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - hi);
}

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == EMER_ACT_RELU) return v > 0.0f ? v : 0.0f;
    if (act == EMER_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
    return v;
}
__device__ __forceinline__ float act_bwd(float g, float y, int act) {
    if (act == EMER_ACT_RELU) return y > 0.0f ? g : 0.0f;
    if (act == EMER_ACT_SIGMOID) return g * (y * (1.0f - y));
    return g;
}

struct Params {
    const float* a;       // fwd: X [n, lda]      bwd: dY [n, lda]
    int64_t lda;
    const float* yact;    // bwd: Y (for act'), may be null when act == none
    int64_t ldy;
    const float* w;       // [n_out, k] row-major
    const float* bias;    // fwd only
    float* c;             // fwd: Y [n, ldc]      bwd: dX [n, ldc]
    int64_t ldc;
    int64_t n;            // rows
    int k, n_out;         // layer widths
    int kred;             // reduction width: fwd k, bwd n_out
    int ncols;            // output width:    fwd n_out, bwd k
    int kred_pad;         // multiple of 8
    int n_pad;            // multiple of 16, <= 256
    int act;
    int accumulate;       // bwd: dX += result
    int tmem_cols;        // power of two >= n_pad, >= 32
};

template <bool BWD>
__global__ void __launch_bounds__(128) tc_linear_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [B_hi | B_lo | A ring: STAGES x (hi, lo) | barriers]
    const int b_panel = p.n_pad * 16;
    const int b_bytes = (p.kred_pad / 4) * b_panel;
    uint8_t* b_hi = smem;
    uint8_t* b_lo = smem + b_bytes;
    uint8_t* a_ring = smem + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + STAGES * 2 * A_STAGE);
    uint64_t* empty_bar = bars;                 // [STAGES]
    uint64_t* accum_bar = bars + STAGES;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + STAGES + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&empty_bar[s], 1);
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- resident B operand: hi/lo panels of W (fwd: B[n][k] = W[n,k]; bwd: B[n=k_in][r=o] = W[o,k_in])
    {
        const int total = p.n_pad * p.kred_pad;
        for (int e = tid; e < total; e += 128) {
            const int r = e % p.kred_pad;      // reduction index
            const int nn = e / p.kred_pad;     // B row (output column of the GEMM)
            float v = 0.0f;
            if (r < p.kred && nn < p.ncols) v = BWD ? __ldg(p.w + (int64_t)r * p.k + nn) : __ldg(p.w + (int64_t)nn * p.k + r);
            float hi, lo;
            split(v, hi, lo);
            const int off = (r >> 2) * b_panel + nn * 16 + (r & 3) * 4;
            *reinterpret_cast<float*>(b_hi + off) = hi;
            *reinterpret_cast<float*>(b_lo + off) = lo;
        }
    }
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = make_idesc(128, p.n_pad);
    const int n_chunks = (p.kred_pad + CHUNK - 1) / CHUNK;
    const bool vec_a = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0) &&
                       (!BWD || p.act == EMER_ACT_NONE || ((p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.yact) & 15) == 0)));

    uint32_t stage_use[STAGES] = {0, 0};      // how many times each ring stage has been filled
    uint32_t tiles_done = 0;
    const int64_t n_tiles = (p.n + ROWS - 1) / ROWS;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * ROWS;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c & 1;
            // the MMAs that last read this stage must have completed
            if (stage_use[s] > 0) mbar_wait(&empty_bar[s], (stage_use[s] - 1) & 1);
            uint8_t* a_hi = a_ring + (s * 2) * A_STAGE;
            uint8_t* a_lo = a_hi + A_STAGE;
            const int k0 = c * CHUNK;
            // 128 rows x 8 float4: thread t -> (row = i*16 + t/8, quad = t%8): a row's 128 B are read by 8
            // consecutive threads (coalesced) and land in 8 different panels (conflict-free, padded LBO)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 16 + (tid >> 3);
                const int q = tid & 7;
                const int kk = k0 + q * 4;
                const int64_t row = row0 + r;
                float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (row < p.n && kk < p.kred) {
                    if (vec_a && kk + 3 < p.lda) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(p.a + row * p.lda + kk));
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        if (BWD && p.act != EMER_ACT_NONE) {
                            const float4 y = __ldg(reinterpret_cast<const float4*>(p.yact + row * p.ldy + kk));
                            v[0] = act_bwd(v[0], y.x, p.act); v[1] = act_bwd(v[1], y.y, p.act);
                            v[2] = act_bwd(v[2], y.z, p.act); v[3] = act_bwd(v[3], y.w, p.act);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (kk + j >= p.kred) v[j] = 0.0f;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (kk + j < p.kred) {
                                float t = __ldg(p.a + row * p.lda + kk + j);
                                if (BWD && p.act != EMER_ACT_NONE) t = act_bwd(t, __ldg(p.yact + row * p.ldy + kk + j), p.act);
                                v[j] = t;
                            }
                        }
                    }
                }
                float4 h, l;
                split(v[0], h.x, l.x); split(v[1], h.y, l.y); split(v[2], h.z, l.z); split(v[3], h.w, l.w);
                *reinterpret_cast<float4*>(a_hi + q * A_PANEL + r * 16) = h;
                *reinterpret_cast<float4*>(a_lo + q * A_PANEL + r * 16) = l;
            }
            fence_async_proxy();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const int ksteps = min(CHUNK, p.kred_pad - k0) / 8;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t a_off = (uint32_t)(ks * 2) * A_PANEL;
                    const uint32_t b_off = (uint32_t)((k0 >> 2) + ks * 2) * b_panel;
                    const uint64_t da_hi = make_desc(smem_u32(a_hi) + a_off, A_PANEL, 128);
                    const uint64_t da_lo = make_desc(smem_u32(a_lo) + a_off, A_PANEL, 128);
                    const uint64_t db_hi = make_desc(smem_u32(b_hi) + b_off, b_panel, 128);
                    const uint64_t db_lo = make_desc(smem_u32(b_lo) + b_off, b_panel, 128);
                    const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
                    mma_tf32(tmem_base, da_lo, db_hi, idesc, first);
                    mma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                    mma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
                }
                tc_commit(&empty_bar[s]);                  // stage reusable once these MMAs retire
                if (c == n_chunks - 1) tc_commit(accum_bar);   // ... and the accumulator is complete
            }
            stage_use[s]++;
        }
        // ---- epilogue: TMEM -> registers -> global.  thread t = row t = TMEM lane t
        mbar_wait(accum_bar, tiles_done & 1);
        tiles_done++;
        tc_fence_after();
        const int64_t row = row0 + tid;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool vec_c = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
        for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(lane_addr + (uint32_t)c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < p.n) {
                float* out = p.c + row * p.ldc + c0;
#pragma unroll
                for (int j0 = 0; j0 < 16; j0 += 4) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = c0 + j0 + j;
                        float v = __uint_as_float(r[j0 + j]);
                        if (!BWD) {
                            if (p.bias && col < p.ncols) v += __ldg(p.bias + col);
                            v = act_fwd(v, p.act);
                        }
                        o[j] = v;
                    }
                    if (vec_c && c0 + j0 + 3 < p.ncols) {
                        float4* dst = reinterpret_cast<float4*>(out + j0);
                        if (BWD && p.accumulate) {
                            const float4 old = *dst;
                            o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
                        }
                        *dst = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (c0 + j0 + j < p.ncols) {
                                if (BWD && p.accumulate) o[j] += out[j0 + j];
                                out[j0 + j] = o[j];
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();      // TMEM reads done before the next tile's first MMA overwrites the accumulator
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static size_t smem_bytes(const Params& p) {
    return (size_t)2 * (p.kred_pad / 4) * p.n_pad * 16 + (size_t)STAGES * 2 * A_STAGE + (STAGES + 1) * 8 + 16;
}

template <bool BWD>
static int launch(Params& p, cudaStream_t st, const char* what) {
    p.kred = BWD ? p.n_out : p.k;
    p.ncols = BWD ? p.k : p.n_out;
    p.kred_pad = round_up(p.kred, 8);
    p.n_pad = round_up(p.ncols, 16);
    EMER_REQUIRE(p.n_pad <= 256, "%s: output width %d exceeds one MMA (256)", what, p.ncols);
    p.tmem_cols = 32;
    while (p.tmem_cols < p.n_pad) p.tmem_cols *= 2;
    const size_t smem = smem_bytes(p);
    EMER_REQUIRE(smem <= 227 * 1024, "%s: layer %dx%d needs %zu B of shared memory", what, p.k, p.n_out, smem);
    static size_t configured[2] = {0, 0};
    if (smem > configured[BWD]) {
        cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute(%zu): %s", what, smem, cudaGetErrorString(e));
            return -2;
        }
        configured[BWD] = smem;
    }
    const int64_t n_tiles = ceil_div(p.n, ROWS);
    const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
    int64_t grid = 148 * ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    tc_linear_kernel<BWD><<<(unsigned)grid, 128, smem, st>>>(p);
    return check_launch(what);
}

}  // namespace tc
}  // namespace emer

using namespace emer;

extern "C" int emer_linear_tc_fwd(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                                  int64_t n, int k, int n_out, int act, void* stream) {
    if (n == 0) return 0;
    EMER_REQUIRE(x && w && y, "emer_linear_tc_fwd: NULL pointer");
    EMER_REQUIRE(k > 0 && n_out > 0 && ldx >= k && ldy >= n_out, "emer_linear_tc_fwd: bad shape k=%d n_out=%d", k, n_out);
    tc::Params p{};
    p.a = x; p.lda = ldx; p.yact = nullptr; p.ldy = 0; p.w = w; p.bias = b; p.c = y; p.ldc = ldy;
    p.n = n; p.k = k; p.n_out = n_out; p.act = act; p.accumulate = 0;
    return tc::launch<false>(p, (cudaStream_t)stream, "emer_linear_tc_fwd");
}

extern "C" int emer_linear_tc_bwd_data(const float* dy, int64_t lddy, const float* y, int64_t ldy, int act,
                                       const float* w, float* dx, int64_t lddx, int64_t n, int k, int n_out,
                                       int accumulate, void* stream) {
    if (n == 0) return 0;
    EMER_REQUIRE(dy && w && dx, "emer_linear_tc_bwd_data: NULL pointer");
    EMER_REQUIRE(act == EMER_ACT_NONE || y, "emer_linear_tc_bwd_data: activation needs the stored output");
    tc::Params p{};
    p.a = dy; p.lda = lddy; p.yact = y; p.ldy = ldy; p.w = w; p.bias = nullptr; p.c = dx; p.ldc = lddx;
    p.n = n; p.k = k; p.n_out = n_out; p.act = act; p.accumulate = accumulate;
    return tc::launch<true>(p, (cudaStream_t)stream, "emer_linear_tc_bwd_data");
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  dW^T[k, n_out] += X^T dZ,  db += sum_rows dZ.
//
// The reduction runs over the ROWS, so both operands are "MN-major": the chunk-major tile layout
// above (offset(r, f) = (f/4)*PANEL + r*16 + (f%4)*4) is also the canonical MN-major SWIZZLE_NONE
// UMMA layout with SBO = PANEL (next 4 features) and 8-row core matrices 128 B apart along the
// reduction (one 8-row core matrix per tf32 k-step).  A = X^T (M = a 128-feature block of k),
// B = dZ^T (N = n_out).  Each persistent CTA keeps dW^T in TMEM (one 128-lane block per 128
// input features) across all of its 64-row tiles and flushes once with atomics.
namespace emer {
namespace tcw {

using namespace emer::tc;

constexpr int WROWS = 64;                     // rows per tile (8 tf32 k-steps)
constexpr int W_PANEL = WROWS * 16 + 16;      // 1040 B

struct WParams {
    const float* x;
    int64_t ldx;
    const float* dy;
    int64_t lddy;
    const float* y;
    int64_t ldy;
    float* dw;       // [n_out, k]
    float* db;       // [n_out] or null
    int64_t n;
    int k, n_out;
    int k_pad4;      // k rounded to 4 (panels of the X tile)
    int n_pad;       // n_out rounded to 16
    int m_blocks;    // ceil(k / 128)
    int x_panels_alloc;   // panels reserved for X (>= 32 * m_blocks so an M=128 operand never leaves smem)
    int act;
    int tmem_cols;
};

__global__ void __launch_bounds__(128) tc_wgrad_kernel(const WParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int x_bytes = p.x_panels_alloc * W_PANEL;
    const int z_panels = p.n_pad / 4;
    const int z_bytes = z_panels * W_PANEL;
    uint8_t* x_hi = smem;
    uint8_t* x_lo = x_hi + x_bytes;
    uint8_t* z_hi = x_lo + x_bytes;
    uint8_t* z_lo = z_hi + z_bytes;
    uint64_t* done_bar = reinterpret_cast<uint64_t*>(z_lo + z_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid == 0) {
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the operand buffers once: padding panels / rows must never hold NaN bit patterns
    for (int i = tid * 16; i < 2 * x_bytes + 2 * z_bytes; i += 128 * 16)
        *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // a=b=tf32, c=f32, A and B MN-major (bits 15, 16)
    const uint32_t idesc = make_idesc(128, p.n_pad) | (1u << 15) | (1u << 16);

    const int64_t n_tiles = (p.n + WROWS - 1) / WROWS;
    uint32_t tiles_done = 0;
    float bsum = 0.0f;        // thread t < n_out owns db[t]
    const int x_quads = p.k_pad4 / 4;
    const bool vec_x = (p.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    const bool vec_z = (p.lddy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dy) & 15) == 0) &&
                       (p.act == EMER_ACT_NONE || ((p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0)));

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * WROWS;
        // the previous tile's MMAs must have finished reading the operand buffers
        if (tiles_done > 0) mbar_wait(done_bar, (tiles_done - 1) & 1);
        // ---- X tile: 64 rows x x_quads float4
        for (int e = tid; e < WROWS * x_quads; e += 128) {
            const int r = e / x_quads, q = e % x_quads;
            const int64_t row = row0 + r;
            const int kk = q * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (row < p.n) {
                if (vec_x && kk + 3 < p.ldx) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p.x + row * p.ldx + kk));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (kk + j >= p.k) v[j] = 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (kk + j < p.k) v[j] = __ldg(p.x + row * p.ldx + kk + j);
                }
            }
            float4 h, l;
            split(v[0], h.x, l.x); split(v[1], h.y, l.y); split(v[2], h.z, l.z); split(v[3], h.w, l.w);
            *reinterpret_cast<float4*>(x_hi + q * W_PANEL + r * 16) = h;
            *reinterpret_cast<float4*>(x_lo + q * W_PANEL + r * 16) = l;
        }
        // ---- dZ tile: 64 rows x z_panels float4, dZ = dY * act'(Y)
        for (int e = tid; e < WROWS * z_panels; e += 128) {
            const int r = e / z_panels, q = e % z_panels;
            const int64_t row = row0 + r;
            const int oo = q * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (row < p.n && oo < p.n_out) {
                if (vec_z && oo + 3 < p.lddy) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p.dy + row * p.lddy + oo));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    if (p.act != EMER_ACT_NONE) {
                        const float4 yy = __ldg(reinterpret_cast<const float4*>(p.y + row * p.ldy + oo));
                        v[0] = act_bwd(v[0], yy.x, p.act); v[1] = act_bwd(v[1], yy.y, p.act);
                        v[2] = act_bwd(v[2], yy.z, p.act); v[3] = act_bwd(v[3], yy.w, p.act);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (oo + j >= p.n_out) v[j] = 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (oo + j < p.n_out) {
                            float t = __ldg(p.dy + row * p.lddy + oo + j);
                            if (p.act != EMER_ACT_NONE) t = act_bwd(t, __ldg(p.y + row * p.ldy + oo + j), p.act);
                            v[j] = t;
                        }
                    }
                }
            }
            float4 h, l;
            split(v[0], h.x, l.x); split(v[1], h.y, l.y); split(v[2], h.z, l.z); split(v[3], h.w, l.w);
            *reinterpret_cast<float4*>(z_hi + q * W_PANEL + r * 16) = h;
            *reinterpret_cast<float4*>(z_lo + q * W_PANEL + r * 16) = l;
        }
        fence_async_proxy();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int mb = 0; mb < p.m_blocks; ++mb) {
                const uint32_t d_addr = tmem_base + (uint32_t)(mb * p.n_pad);
                const uint32_t xa = (uint32_t)(mb * 32) * W_PANEL;       // 128 features = 32 panels
                for (int ks = 0; ks < WROWS / 8; ++ks) {
                    const uint32_t roff = (uint32_t)ks * 128;            // next 8 rows
                    const uint64_t da_hi = make_desc(smem_u32(x_hi) + xa + roff, 128, W_PANEL);
                    const uint64_t da_lo = make_desc(smem_u32(x_lo) + xa + roff, 128, W_PANEL);
                    const uint64_t db_hi = make_desc(smem_u32(z_hi) + roff, 128, W_PANEL);
                    const uint64_t db_lo = make_desc(smem_u32(z_lo) + roff, 128, W_PANEL);
                    const uint32_t acc = (tiles_done == 0 && ks == 0) ? 0u : 1u;
                    mma_tf32(d_addr, da_lo, db_hi, idesc, acc);
                    mma_tf32(d_addr, da_hi, db_lo, idesc, 1u);
                    mma_tf32(d_addr, da_hi, db_hi, idesc, 1u);
                }
            }
            tc_commit(done_bar);
        }
        // bias gradient from the staged dZ tile (hi + lo = the fp32 value to ~2^-21)
        if (p.db && tid < p.n_out) {
            const uint8_t* zh = z_hi + (tid >> 2) * W_PANEL + (tid & 3) * 4;
            const uint8_t* zl = z_lo + (tid >> 2) * W_PANEL + (tid & 3) * 4;
#pragma unroll 8
            for (int r = 0; r < WROWS; ++r)
                bsum += *reinterpret_cast<const float*>(zh + r * 16) + *reinterpret_cast<const float*>(zl + r * 16);
        }
        tiles_done++;
    }
    if (tiles_done > 0) {
        mbar_wait(done_bar, (tiles_done - 1) & 1);
        tc_fence_after();
        // flush: lane f of block mb holds dW^T[mb*128 + f, :]
        for (int mb = 0; mb < p.m_blocks; ++mb) {
            const int f = mb * 128 + tid;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mb * p.n_pad);
            for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(lane_addr + (uint32_t)c0)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (f < p.k) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int o = c0 + j;
                        if (o < p.n_out) atomicAdd(p.dw + (int64_t)o * p.k + f, __uint_as_float(r[j]));
                    }
                }
            }
        }
        if (p.db && tid < p.n_out) atomicAdd(p.db + tid, bsum);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

}  // namespace tcw
}  // namespace emer

extern "C" int emer_linear_tc_bwd_weight(const float* x, int64_t ldx, const float* dy, int64_t lddy, const float* y,
                                         int64_t ldy, int act, float* dw, float* db, int64_t n, int k, int n_out,
                                         void* stream) {
    using namespace emer::tcw;
    if (n == 0) return 0;
    EMER_REQUIRE(x && dy && dw, "emer_linear_tc_bwd_weight: NULL pointer");
    EMER_REQUIRE(act == EMER_ACT_NONE || y, "emer_linear_tc_bwd_weight: activation needs the stored output");
    EMER_REQUIRE(n_out <= 128 && k <= 256, "emer_linear_tc_bwd_weight: widths k=%d n_out=%d out of range", k, n_out);
    WParams p{};
    p.x = x; p.ldx = ldx; p.dy = dy; p.lddy = lddy; p.y = y; p.ldy = ldy; p.dw = dw; p.db = db;
    p.n = n; p.k = k; p.n_out = n_out; p.act = act;
    p.k_pad4 = (k + 3) / 4 * 4;
    p.n_pad = (n_out + 15) / 16 * 16;
    p.m_blocks = (k + 127) / 128;
    p.x_panels_alloc = 32 * p.m_blocks;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.m_blocks * p.n_pad) p.tmem_cols *= 2;
    const size_t smem = (size_t)2 * p.x_panels_alloc * W_PANEL + (size_t)2 * (p.n_pad / 4) * W_PANEL + 8 + 16;
    EMER_REQUIRE(smem <= 227 * 1024, "emer_linear_tc_bwd_weight: %zu B of shared memory", smem);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            emer::set_error("emer_linear_tc_bwd_weight: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return -2;
        }
        configured = smem;
    }
    const int64_t n_tiles = emer::ceil_div(n, WROWS);
    const int ctas_per_sm = smem <= 75 * 1024 ? 3 : (smem <= 113 * 1024 ? 2 : 1);
    int64_t grid = 148 * ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    tc_wgrad_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(p);
    return emer::check_launch("emer_linear_tc_bwd_weight");
}
