// Probe: does tcgen05.mma.kind::tf32 take MN-major shared-memory operands with SWIZZLE_128B?  (sm_100a)
//
// Why: a weight gradient dW^T[k, o] = sum_rows X[row, k] dZ[row, o] reduces over the ROWS, so both operands are
// "MN-major" in their natural row-major layout (consecutive memory runs along the feature = M / N index).  With
// K-major-only operands the kernel has to transpose while staging (linear_tc.cu: 4 rows x 1 feature gathers).  If the
// MN-major 128B-swizzled layout works, a row-major tile of 32 features x 8 rows (128-byte rows, 16-byte chunks XOR-ed
// with row % 8 -- what a TMA box load with SWIZZLE_128B produces) is an operand as it lies, and the tf32 hi / lo split
// becomes elementwise.
//
// Canonical layout tested (cute: Swizzle<3,4,3> o ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)), T = 4 tf32 per 16 bytes):
//   element (mn, k):  (mn / 32) * LBO + (k / 8) * SBO + (k % 8) * 128 + (((mn % 32) / 4) ^ (k % 8)) * 16 + (mn % 4) * 4
//   nvcc -gencode arch=compute_100a,code=sm_100a -o mn_probe mn_probe.cu && ./mn_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define M 128
#define N 64
#define K 32

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;          // 0 none, 2 SWIZZLE_128B
    return d;
}
__host__ __device__ inline float aval(int m, int k) { return (float)((m % 7) - 3) + 0.25f * (float)(k % 8) + (float)(k / 8); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n % 5) - 2) + 0.5f * (float)(k % 8) - (float)(k / 8); }

struct Variant {
    int layout;                         // 2: SWIZZLE_128B (16-byte chunks ^ (k % 8), 8-deep k atom)
                                        // 1: SWIZZLE_128B_BASE32B (32-byte chunks ^ (k % 4), 4-deep k atom; cute's
                                        //    Layout_MN_SW128_32B_Atom, the atom CUTLASS picks for 32-bit MN-major operands)
    int a_lbo, a_sbo, b_lbo, b_sbo;     // descriptor fields (bytes)
    int a_mstride, a_kstride;           // where the probe PLACES the 32-wide mn blocks / 8-deep k groups (bytes)
    int b_mstride, b_kstride;
    int a_kadv, b_kadv;                 // descriptor start-address advance per 8-deep k step (bytes)
};

__global__ void probe(const Variant v, float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + 64 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 1024) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
    __syncthreads();
    auto place = [&](int mn, int k, int mstride, int kstride) {
        if (v.layout == 2)
            return (mn / 32) * mstride + (k / 8) * kstride + (k % 8) * 128 + ((((mn % 32) / 4) ^ (k % 8)) * 16) + (mn % 4) * 4;
        return (mn / 32) * mstride + (k / 4) * kstride + (k % 4) * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4;
    };
    for (int e = tid; e < M * K; e += 128) {
        const int m = e % M, k = e / M;
        *reinterpret_cast<float*>(a_s + place(m, k, v.a_mstride, v.a_kstride)) = aval(m, k);
    }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e % N, k = e / N;
        *reinterpret_cast<float*>(b_s + place(n, k, v.b_mstride, v.b_kstride)) = bval(n, k);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        idesc |= (1u << 15) | (1u << 16);                     // A and B MN-major
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = make_desc(smem_u32(a_s) + ks * v.a_kadv, v.a_lbo, v.a_sbo, v.layout);
            const uint64_t db = make_desc(smem_u32(b_s) + ks * v.b_kadv, v.b_lbo, v.b_sbo, v.layout);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db),
                         "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, M * N * 4);
    static float h[M * N];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    // placements: mn blocks of 32 features; k groups of 8 rows (1024 B each).
    //   "k-inner":  k groups contiguous (1024 B apart), mn blocks after all k groups      (A: 4 KB apart for K = 32)
    //   "mn-inner": mn blocks contiguous (1024 B apart), k groups after all mn blocks
    const int KG = K / 8;
    const int MB_A = M / 32, MB_B = N / 32;       // 32-wide mn blocks
    struct { const char* name; Variant v; } tests[] = {
        // SWIZZLE_128B, 8-deep k atoms of 1024 B
        {"SW128        mn-inner LBO=1024 SBO=mnblk*1024          ", {2, 1024, MB_A * 1024, 1024, MB_B * 1024, 1024, MB_A * 1024, 1024, MB_B * 1024, MB_A * 1024, MB_B * 1024}},
        {"SW128        mn-inner LBO/SBO swapped                  ", {2, MB_A * 1024, 1024, MB_B * 1024, 1024, 1024, MB_A * 1024, 1024, MB_B * 1024, MB_A * 1024, MB_B * 1024}},
        {"SW128        k-inner  LBO=KG*1024 SBO=1024             ", {2, KG * 1024, 1024, KG * 1024, 1024, KG * 1024, 1024, KG * 1024, 1024, 1024, 1024}},
        {"SW128        k-inner  LBO/SBO swapped                  ", {2, 1024, KG * 1024, 1024, KG * 1024, KG * 1024, 1024, KG * 1024, 1024, 1024, 1024}},
        // SWIZZLE_128B_BASE32B, 4-deep k atoms of 512 B; one MMA (k = 8) spans two atoms
        {"SW128_BASE32 mn-inner LBO=512 SBO=mnblk*512 adv=2*SBO  ", {1, 512, MB_A * 512, 512, MB_B * 512, 512, MB_A * 512, 512, MB_B * 512, 2 * MB_A * 512, 2 * MB_B * 512}},
        {"SW128_BASE32 mn-inner LBO/SBO swapped                  ", {1, MB_A * 512, 512, MB_B * 512, 512, 512, MB_A * 512, 512, MB_B * 512, 2 * MB_A * 512, 2 * MB_B * 512}},
        {"SW128_BASE32 k-inner  LBO=2KG*512 SBO=512   adv=1024   ", {1, 2 * KG * 512, 512, 2 * KG * 512, 512, 2 * KG * 512, 512, 2 * KG * 512, 512, 1024, 1024}},
        {"SW128_BASE32 k-inner  LBO/SBO swapped                  ", {1, 512, 2 * KG * 512, 512, 2 * KG * 512, 2 * KG * 512, 512, 2 * KG * 512, 512, 1024, 1024}},
    };
    for (auto& t : tests) {
        cudaMemset(d_out, 0xff, M * N * 4);
        probe<<<1, 128, 128 * 1024>>>(t.v, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s : CUDA error %s\n", t.name, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d_out, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double want = 0; for (int k = 0; k < K; ++k) want += (double)aval(m, k) * bval(n, k);
            double err = fabs(want - h[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) bad++;
        }
        printf("%s : %s  max|err| %.4g  bad %d/%d  D[0,0]=%g D[33,40]=%g D[127,63]=%g\n", t.name, bad == 0 ? "MATCH" : "wrong", maxerr,
               bad, M * N, h[0], h[33 * N + 40], h[127 * N + 63]);
    }
    return 0;
}
