// Probe of tcgen05.mma operand layouts on real hardware (sm_100a).  Builds one 128x64xK tf32 MMA
// with exactly-representable inputs for several (smem layout, descriptor, major-ness) variants and
// reports which ones reproduce D = A * B^T.  Used to pin the MN-major (transposed-operand) layout
// of the weight-gradient kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
// Add -DPROBE_ELECT=1 to issue the MMAs from an elect.sync lane (cycles per MMA without the waterfall loops).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define M 128
#define N 64

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;          // LayoutType::SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ float aval(int m, int k) { return (float)((m % 7) - 3) + 0.25f * (float)(k % 8) + (float)(k / 8); }
__device__ float bval(int n, int k) { return (float)((n % 5) - 2) + 0.5f * (float)(k % 8) - (float)(k / 8); }

struct Variant {
    int sw128;                 // 1: K-major SWIZZLE_128B operands (rows of 128 B, 16-B chunks XOR (row%8))
    int a_major, b_major;      // 0 = K-major, 1 = MN-major
    int a_lbo, a_sbo, b_lbo, b_sbo;
    int a_kstep, b_kstep;      // byte advance of the start address per 8-wide k step
    int ksteps;
};

// chunk layout, element (r = row in MN, k): (k/4)*panel + r*16 + (k%4)*4      [K-major use]
// same memory viewed MN-major: element (mn = f, k = r): (f/4)*panel + r*16 + (f%4)*4
// The MMA-issuing lane.  PROBE_ELECT=0: thread 0 by index (the compiler then wraps every tcgen05.mma in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop -- the 100-165 cycles per MMA first measured with this probe
// are that loop, see profiles/r1_sass_mma_issue.md).  PROBE_ELECT=1: elect.sync, descriptors in uniform registers.
#ifndef PROBE_ELECT
#define PROBE_ELECT 0
#endif
__device__ __forceinline__ bool issue_lane(int tid) {
#if PROBE_ELECT
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
#else
    return tid == 0;
#endif
}
__global__ void probe(const Variant v, float* out, int a_panel, int b_panel, int K, long long* cycles, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + 96 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (192 * 1024) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
    __syncthreads();
    if (v.sw128) {
        for (int e = tid; e < M * K; e += 128) { int m = e / K, k = e % K;
            *reinterpret_cast<float*>(a_s + m * 128 + ((((k / 4) ^ (m % 8)) & 7) * 16) + (k % 4) * 4) = aval(m, k); }
        for (int e = tid; e < N * K; e += 128) { int n = e / K, k = e % K;
            *reinterpret_cast<float*>(b_s + n * 128 + ((((k / 4) ^ (n % 8)) & 7) * 16) + (k % 4) * 4) = bval(n, k); }
    } else {
    if (v.a_major == 0) {
        for (int e = tid; e < M * K; e += 128) { int m = e / K, k = e % K;
            *reinterpret_cast<float*>(a_s + (k / 4) * a_panel + m * 16 + (k % 4) * 4) = aval(m, k); }
    } else {   // MN-major: "row" of the panel = k (reduction index), feature = m
        for (int e = tid; e < M * K; e += 128) { int m = e / K, k = e % K;
            *reinterpret_cast<float*>(a_s + (m / 4) * a_panel + k * 16 + (m % 4) * 4) = aval(m, k); }
    }
    if (v.b_major == 0) {
        for (int e = tid; e < N * K; e += 128) { int n = e / K, k = e % K;
            *reinterpret_cast<float*>(b_s + (k / 4) * b_panel + n * 16 + (k % 4) * 4) = bval(n, k); }
    } else {
        for (int e = tid; e < N * K; e += 128) { int n = e / K, k = e % K;
            *reinterpret_cast<float*>(b_s + (n / 4) * b_panel + k * 16 + (n % 4) * 4) = bval(n, k); }
    }
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
    if (warp == 0 && issue_lane(tid)) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        idesc |= (uint32_t)v.a_major << 15;
        idesc |= (uint32_t)v.b_major << 16;
        long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
        for (int ks = 0; ks < v.ksteps; ++ks) {
            uint64_t da = v.sw128 ? make_desc_sw128(smem_u32(a_s) + ks * v.a_kstep, v.a_lbo, v.a_sbo)
                                  : make_desc(smem_u32(a_s) + ks * v.a_kstep, v.a_lbo, v.a_sbo);
            uint64_t db = v.sw128 ? make_desc_sw128(smem_u32(b_s) + ks * v.b_kstep, v.b_lbo, v.b_sbo)
                                  : make_desc(smem_u32(b_s) + ks * v.b_kstep, v.b_lbo, v.b_sbo);
            uint32_t acc = (ks > 0) || (rep > 0 && rep < reps - 1 ? 0 : 0) ;
            if (rep > 0) acc = (ks > 0);     // every repetition recomputes the same D (first k-step overwrites)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db),
                         "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok2 = 0;
        while (!ok2) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok2) : "r"(smem_u32(&bar)) : "memory");
        }
        *cycles = clock64() - t0;
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

static float ha(int m, int k) { return (float)((m % 7) - 3) + 0.25f * (float)(k % 8) + (float)(k / 8); }
static float hb(int n, int k) { return (float)((n % 5) - 2) + 0.5f * (float)(k % 8) - (float)(k / 8); }

int main() {
    float* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, M * N * 4);
    cudaMalloc(&d_cyc, 8);
    static float h[M * N];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    const int PA = 2064, PB = 1040;     // panel strides used by the library kernels (padded)
    struct { const char* name; Variant v; int a_panel, b_panel, K; } tests[] = {
        // known-good K-major reference (forward kernel layout): a_panel = 128*16+16
        {"K/K  NONE  lbo=panel sbo=128                ", {0, 0, 0, PA, 128, 1040, 128, 2 * PA, 2 * 1040, 2}, PA, 1040, 16},
        {"K/K  SW128 lbo=16 sbo=1024 kstep=32B        ", {1, 0, 0, 16, 1024, 16, 1024, 32, 32, 2}, 0, 0, 16},
        {"K/K  SW128 lbo=0  sbo=1024 kstep=32B        ", {1, 0, 0, 0, 1024, 0, 1024, 32, 32, 2}, 0, 0, 16},
        {"K/K  SW128 lbo=16 sbo=1024 kstep=32B  K=32  ", {1, 0, 0, 16, 1024, 16, 1024, 32, 32, 4}, 0, 0, 32},
        // MN-major candidates: reduction rows at 16 B inside a panel, 4 features per panel
        {"MN/MN lbo=128   sbo=panel (cute INTERLEAVE)  ", {0, 1, 1, 128, PB, 128, PB, 128, 128, 2}, PB, PB, 16},
        {"MN/MN lbo=panel sbo=128                      ", {0, 1, 1, PB, 128, PB, 128, 128, 128, 2}, PB, PB, 16},
        {"MN/K  lbo=128   sbo=panel | K lbo=panel      ", {0, 1, 0, 128, PB, 1040, 128, 128, 2 * 1040, 2}, PB, 1040, 16},
        {"MN/K  lbo=panel sbo=128   | K lbo=panel      ", {0, 1, 0, PB, 128, 1040, 128, 128, 2 * 1040, 2}, PB, 1040, 16},
        {"K/MN  K lbo=panel | lbo=128 sbo=panel        ", {0, 0, 1, PA, 128, 128, PB, 2 * PA, 128, 2}, PA, PB, 16},
        {"K/MN  K lbo=panel | lbo=panel sbo=128        ", {0, 0, 1, PA, 128, PB, 128, 2 * PA, 128, 2}, PA, PB, 16},
    };
    for (auto& t : tests) {
        cudaMemset(d_out, 0xff, M * N * 4);
        probe<<<1, 128, 192 * 1024>>>(t.v, d_out, t.a_panel, t.b_panel, t.K, d_cyc, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s : CUDA error %s\n", t.name, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d_out, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double want = 0; for (int k = 0; k < t.K; ++k) want += (double)ha(m, k) * hb(n, k);
            double err = fabs(want - h[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) bad++;
        }
        double w00 = 0, w12 = 0; for (int k = 0; k < t.K; ++k) { w00 += (double)ha(0, k) * hb(0, k); w12 += (double)ha(1, k) * hb(2, k); }
        long long c1 = 0, c2 = 0;
        cudaMemcpy(&c1, d_cyc, 8, cudaMemcpyDeviceToHost);
        probe<<<1, 128, 192 * 1024>>>(t.v, d_out, t.a_panel, t.b_panel, t.K, d_cyc, 257);
        cudaDeviceSynchronize();
        cudaMemcpy(&c2, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("%s : %s  max|err| %.4g  bad %d/%d   D[0,0]=%g (want %g)  D[1,2]=%g (want %g)   cycles/MMA %.1f\n", t.name,
               bad == 0 ? "MATCH" : "wrong", maxerr, bad, M * N, h[0], w00, h[1 * N + 2], w12,
               (double)(c2 - c1) / (256.0 * t.v.ksteps));
    }
    return 0;
}
