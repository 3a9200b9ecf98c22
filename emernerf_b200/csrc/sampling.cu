// Inverse-CDF resampling of ray intervals + the s->t warp.
//
// Replaces nerfacc.pdf.importance_sampling (the one nerfacc CUDA kernel the reference hits by
// default, third_party/nerfacc_prop_net.py:153,172) and _transform_stot (:317-339).
// Integer work (upper-bound bin per output edge) is bit-exact against the oracle; the fp32
// arithmetic is written operation by operation in the oracle's order (library built with
// -fmad=false), so the produced s/t edges are bit-identical for identical CDFs.
//
// Layout: vals, cdfs [R, m1] row-major; outputs [R, n+1].  One thread per output edge;
// a ray's CDF row (<= 129 floats) is read through L1.  HBM-bound and tiny:
// (2*m1 + 2*(n+1)) * 4 bytes per ray.
#include "common.cuh"

namespace emer {

__device__ __forceinline__ float s_to_t(float s, float s_min, float s_max, int kind) {
    // icontract(s * s_max + (1 - s) * s_min)
    const float v = s * s_max + (1.0f - s) * s_min;
    switch (kind) {
        case EMER_STOT_UNIFORM: return v;
        case EMER_STOT_LINDISP: return 1.0f / v;
        case EMER_STOT_SQRT: return v * v;
        case EMER_STOT_LOG: return expf(v);
        // torch evaluates `200 / x` as reciprocal(x) * 200 (Tensor.__rtruediv__): two roundings
        case EMER_STOT_UNIFORM_LINDISP: return v < 0.5f ? v * 400.0f : (1.0f / (2.0f - 2.0f * v)) * 200.0f;
        default: return v < 0.5f ? 2.0f * v : 1.0f / (2.0f - 2.0f * v);
    }
}

__global__ void pdf_resample_kernel(const float* __restrict__ vals, const float* __restrict__ cdfs,
                                    int m1, int n, const float* __restrict__ bias, float s_min,
                                    float s_max, int kind, float* __restrict__ out_s,
                                    float* __restrict__ out_t, int32_t* __restrict__ out_bins,
                                    int64_t total) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int64_t ray = tid / (n + 1);
    const int k = (int)(tid - ray * (n + 1));
    const float* c = cdfs + ray * m1;
    const float* v = vals + ray * m1;
    const float u_floor = __ldg(c);
    const float u_ceil = __ldg(c + m1 - 1);
    const float u_step = (u_ceil - u_floor) / (float)n;
    const float b = bias ? __ldg(bias + ray) : 0.5f;
    const float u = u_floor + ((float)k + (b - 0.5f)) * u_step;
    // upper bound: first p in [0, m1] with c[p] > u
    int lo = 0, hi = m1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(c + mid) > u) hi = mid;
        else lo = mid + 1;
    }
    const int p = lo;
    const int p0 = min(max(p - 1, 0), m1 - 1);
    const int p1 = min(max(p, 0), m1 - 1);
    const float u_lo = __ldg(c + p0), u_hi = __ldg(c + p1);
    const float t_lo = __ldg(v + p0), t_hi = __ldg(v + p1);
    const float du = u_hi - u_lo;
    float s;
    if (du < 1e-10f) s = (t_lo + t_hi) * 0.5f;
    else s = (u - u_lo) * ((t_hi - t_lo) / du) + t_lo;
    out_s[tid] = s;
    if (out_t) out_t[tid] = s_to_t(s, s_min, s_max, kind);
    if (out_bins) out_bins[tid] = p;
}

}  // namespace emer

using namespace emer;

extern "C" int emer_pdf_resample(const float* vals, const float* cdfs, int m1, int n, const float* bias,
                                 float s_min, float s_max, int stot_kind, float* out_s, float* out_t,
                                 int32_t* out_bins, int64_t n_rays, void* stream) {
    if (n_rays == 0) return 0;
    EMER_REQUIRE(vals && cdfs && out_s, "emer_pdf_resample: NULL pointer");
    EMER_REQUIRE(m1 >= 2 && n >= 1, "emer_pdf_resample: need m1 >= 2 edges and n >= 1 intervals");
    EMER_REQUIRE(stot_kind >= 0 && stot_kind <= EMER_STOT_UNIFORM_LINDISP_0, "emer_pdf_resample: unknown s->t kind %d",
                 stot_kind);
    const int64_t total = n_rays * (n + 1);
    pdf_resample_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        vals, cdfs, m1, n, bias, s_min, s_max, stot_kind, out_s, out_t, out_bins, total);
    return check_launch("emer_pdf_resample");
}
