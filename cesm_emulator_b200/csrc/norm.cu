// Bandwidth-bound normalisation kernels (fp16 channels-last activations, fp32 math/statistics).
//
//  GroupNorm + FiLM + SiLU (+ residual)      reference: video_net.py:216-227 (Block), :265 (+res)
//  channel LayerNorm (gain only)              reference: video_net.py:78-87
//
// Layout: x[b][p][c], p = (frame, row, col) flattened, c fastest.  Every thread owns one 16-byte
// vector (8 fp16 channels) of a pixel, so global accesses are fully coalesced 128-bit
// transactions; reductions go thread -> shared memory -> one fp32 atomic per (block, slot).
#include "api_common.h"
#include "common.cuh"

namespace cesm {

static constexpr int kNormThreads = 256;

struct Vec8 {
    float v[8];
};
__device__ __forceinline__ Vec8 load8(const h16* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    Vec8 r;
    float2 a = unpack_h2(u.x), b = unpack_h2(u.y), c = unpack_h2(u.z), d = unpack_h2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ void store8(h16* p, const Vec8& r) {
    uint4 u;
    u.x = pack_h2(r.v[0], r.v[1]);
    u.y = pack_h2(r.v[2], r.v[3]);
    u.z = pack_h2(r.v[4], r.v[5]);
    u.w = pack_h2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: sums[b][g] = (sum x, sum x^2) over the group's channels and all pixels.
// grid = (blocks_per_sample, B).  Requires (C/G) % 8 == 0 and 2048 % C == 0.
// Each thread keeps UNR independent 16-byte loads in flight per trip (latency hiding: these
// kernels are pure streams, so bytes in flight per SM is what sets the achieved bandwidth).
// ------------------------------------------------------------------------------------------------
static constexpr int UNR = 4;
static constexpr int PF = 2;   // vectors per tensor per software-pipeline stage in the backward kernels
static constexpr int GN_STAGES = 3;
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
    return u;
}
__device__ __forceinline__ uint4 ld_stream16(const void* p) {  // read-once stream: keep it out of L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kNormThreads)
gn_stats_kernel(const h16* __restrict__ x, float* __restrict__ sums, long long P, int C, int G) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int slot = threadIdx.x % vec_per_pix;       // fixed channel vector of this thread
    const int pix_per_iter = kNormThreads / vec_per_pix;
    const h16* xb = x + (size_t)b * P * C + slot * 8;
    const long long stride = (long long)gridDim.x * pix_per_iter;
    float s = 0.f, ss = 0.f;
    for (long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix; p < P; p += UNR * stride) {
        uint4 u[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const long long pk = p + k * stride;
            u[k] = pk < P ? __ldg(reinterpret_cast<const uint4*>(xb + pk * C)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h2(w[i]);
                s += f.x + f.y;
                ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
            }
        }
    }
    // block reduction per group: shared-memory atomics (a few-way conflict per warp), then one
    // global atomic per (block, group, moment)
    __shared__ float sh[2 * 256];
    for (int i = threadIdx.x; i < 2 * G; i += kNormThreads) sh[i] = 0.f;
    __syncthreads();
    const int g = slot / ((C / G) >> 3);
    atomicAdd(&sh[2 * g], s);
    atomicAdd(&sh[2 * g + 1], ss);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += kNormThreads) atomicAdd(&sums[(size_t)b * G * 2 + i], sh[i]);
}

// out = silu(((x-mean)*rstd*gamma + beta) * (scale+1) + shift) (+ residual)
__global__ void __launch_bounds__(kNormThreads)
gn_apply_fwd_kernel(const h16* __restrict__ x, const float* __restrict__ sums,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ film,  // [B][2C] (scale | shift) or null
                    const h16* __restrict__ residual, h16* __restrict__ out, long long P, int C,
                    int G, float eps) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int slot = threadIdx.x % vec_per_pix;
    const int pix_per_iter = kNormThreads / vec_per_pix;
    const int g = (slot * 8) / (C / G);
    const float cnt = (float)P * (float)(C / G);
    const float mean = sums[((size_t)b * G + g) * 2] / cnt;
    const float var = fmaxf(sums[((size_t)b * G + g) * 2 + 1] / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float A[8], Bc[8];  // u = x*A + Bc
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = slot * 8 + i;
        float ga = gamma[c] * rstd, be = beta[c] - mean * rstd * gamma[c];
        if (film) {
            const float sc = film[(size_t)b * 2 * C + c] + 1.f, sh = film[(size_t)b * 2 * C + C + c];
            ga *= sc;
            be = be * sc + sh;
        }
        A[i] = ga;
        Bc[i] = be;
    }
    float2 A2[4], B2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        A2[i] = make_float2(0.5f * A[2 * i], 0.5f * A[2 * i + 1]);
        B2[i] = make_float2(0.5f * Bc[2 * i], 0.5f * Bc[2 * i + 1]);
    }
    const size_t base = (size_t)b * P * C + slot * 8;
    const long long stride = (long long)gridDim.x * pix_per_iter;
    for (long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix; p < P; p += UNR * stride) {
        uint4 u[UNR], r[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const long long pk = p + k * stride;
            if (pk < P) {
                u[k] = __ldg(reinterpret_cast<const uint4*>(x + base + pk * C));
                if (residual) r[k] = __ldg(reinterpret_cast<const uint4*>(residual + base + pk * C));
            }
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const long long pk = p + k * stride;
            if (pk >= P) break;
            const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
            const uint32_t rw[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // packed pairs: h = u/2 (A2, B2 are the halved coefficients), silu(u) = h (1 + tanh h)
                const float2 h = ffma2(unpack_h2(w[i]), A2[i], B2[i]);
                float2 ov = ffma2(h, tanh2(h), h);
                if (residual) ov = fadd2(ov, unpack_h2(rw[i]));
                o[i] = pack_h2(ov.x, ov.y);
            }
            *reinterpret_cast<uint4*>(out + base + pk * C) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// Backward pass 1: per (b, c) sums over pixels of  [0] du, [1] du*xhat, [2] x
//   with z = xhat*gamma+beta, u = z*(sc+1)+sh, du = dout*silu'(u).  Everything else the backward
//   needs is algebra on these: dz = du*(sc+1), sum dz = (sc+1) T0, sum dz*xhat = (sc+1) T1,
//   sum du*z = gamma T1 + beta T0.
__global__ void __launch_bounds__(kNormThreads, 3)
gn_bwd_reduce_kernel(const h16* __restrict__ x, const h16* __restrict__ dout,
                     const float* __restrict__ sums, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ film, float* __restrict__ csum,
                     long long P, int C, int G, float eps) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int slot = threadIdx.x % vec_per_pix;
    const int pix_per_iter = kNormThreads / vec_per_pix;
    // the first ring stages are requested before anything else: the statistics / coefficient loads below
    // (a chain of dependent global loads, ~2 us) then overlap with them instead of preceding them
    const size_t base = (size_t)b * P * C + slot * 8;
    const long long stride = (long long)gridDim.x * pix_per_iter;
    // thread-private cp.async ring ([stage][vector][thread] x 16 B): trips i+1 and i+2 are in flight while
    // trip i is reduced; a thread reads back only its own copies, so there is no barrier in the loop
    extern __shared__ __align__(16) uint8_t gn_ring[];
    const uint32_t ring = smem_u32(gn_ring) + threadIdx.x * 16u;
    constexpr uint32_t kVec = kNormThreads * 16, kStage = 2 * PF * kVec;
    auto issue = [&](long long p, int stage) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const long long pk = p + k * stride;
            if (pk < P) {
                cp_async16_cg(ring + stage * kStage + (2 * k) * kVec, x + base + pk * C);
                cp_async16_cg(ring + stage * kStage + (2 * k + 1) * kVec, dout + base + pk * C);
            }
        }
        cp_async_commit_group();
    };
    long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix;
#pragma unroll
    for (int s0 = 0; s0 < GN_STAGES - 1; ++s0) issue(p + (long long)s0 * PF * stride, s0);
    const int g = (slot * 8) / (C / G);
    const float cnt = (float)P * (float)(C / G);
    const float mean = sums[((size_t)b * G + g) * 2] / cnt;
    const float var = fmaxf(sums[((size_t)b * G + g) * 2 + 1] / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float A[8], Bc[8];  // u = x*A + Bc
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = slot * 8 + i;
        float ga = gamma[c] * rstd, be = beta[c] - mean * rstd * gamma[c];
        if (film) {
            const float sc = film[(size_t)b * 2 * C + c] + 1.f, sh = film[(size_t)b * 2 * C + C + c];
            ga *= sc;
            be = be * sc + sh;
        }
        A[i] = 0.5f * ga;  // u / 2 (see dsilu_R_from_half)
        Bc[i] = 0.5f * be;
    }
    float2 A2[4], B2[4], acc0[4], acc1[4], acc2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        A2[i] = make_float2(A[2 * i], A[2 * i + 1]);
        B2[i] = make_float2(Bc[2 * i], Bc[2 * i + 1]);
        acc0[i] = acc1[i] = acc2[i] = make_float2(0.f, 0.f);
    }
    int stage = 0;
    for (; p < P; p += PF * stride) {
        {
            int ns = stage + GN_STAGES - 1;
            if (ns >= GN_STAGES) ns -= GN_STAGES;
            issue(p + (long long)(GN_STAGES - 1) * PF * stride, ns);
        }
        cp_async_wait_group<GN_STAGES - 1>();
        uint4 u[PF], d[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            if (p + k * stride < P) {
                u[k] = lds16(ring + stage * kStage + (2 * k) * kVec);
                d[k] = lds16(ring + stage * kStage + (2 * k + 1) * kVec);
            } else {
                u[k] = make_uint4(0, 0, 0, 0);
                d[k] = make_uint4(0, 0, 0, 0);  // dout = 0 -> contributes nothing to T0/T1; x = 0 -> nothing to T2
            }
        }
        if (++stage == GN_STAGES) stage = 0;
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
            const uint32_t dw[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h2(w[i]), dd = unpack_h2(dw[i]);
                // packed pairs; A2, B2 hold u/2 coefficients: h = u/2, t = tanh(h), R = t + h (1 - t^2),
                // du = dout * silu'(u) = (dout/2) (1 + R)
                const float2 h = ffma2(f, A2[i], B2[i]);
                const float2 t = tanh2(h);
                const float2 R = ffma2(h, ffma2(make_float2(-t.x, -t.y), t, make_float2(1.f, 1.f)), t);
                const float2 e = fmul2(dd, make_float2(0.5f, 0.5f));
                const float2 du = ffma2(e, R, e);
                acc0[i] = fadd2(acc0[i], du);
                acc1[i] = ffma2(du, f, acc1[i]);  // sum du*x; xhat folded in below
                acc2[i] = fadd2(acc2[i], f);
            }
        }
    }
    float acc[3][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc[0][2 * i] = acc0[i].x; acc[0][2 * i + 1] = acc0[i].y;
        acc[1][2 * i] = acc1[i].x; acc[1][2 * i + 1] = acc1[i].y;
        acc[2][2 * i] = acc2[i].x; acc[2][2 * i + 1] = acc2[i].y;
    }
    // sum du*xhat = rstd * (sum du*x - mean * sum du)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[1][i] = rstd * (acc[1][i] - mean * acc[0][i]);
    // reduce threads that share a channel slot (stride vec_per_pix) through shared memory
    __shared__ float red[kNormThreads][9];
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[k][i];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += kNormThreads) {  // thread t sums channels t, t+256, ...
            float s = 0.f;
            for (int t = c >> 3; t < kNormThreads; t += vec_per_pix) s += red[t][c & 7];
            atomicAdd(&csum[((size_t)b * C + c) * 3 + k], s);
        }
        __syncthreads();
    }
}

// Backward pass 2: dx = rstd * (gamma*dz - m1 - xhat*m2), m1/m2 = group means of gamma*dz and gamma*dz*xhat.
__global__ void __launch_bounds__(kNormThreads, 3)
gn_bwd_apply_kernel(const h16* __restrict__ x, const h16* __restrict__ dout,
                    const float* __restrict__ sums, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ film,
                    const float* __restrict__ csum, h16* __restrict__ dx, long long P, int C, int G,
                    float eps, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dfilm,
                    float* __restrict__ dcbias) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int slot = threadIdx.x % vec_per_pix;
    const int pix_per_iter = kNormThreads / vec_per_pix;
    const int cpg = C / G;
    const int g = (slot * 8) / cpg;
    const float cnt = (float)P * (float)cpg;
    const float mean = sums[((size_t)b * G + g) * 2] / cnt;
    const float var = fmaxf(sums[((size_t)b * G + g) * 2 + 1] / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float m1 = 0.f, m2 = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        const float sc = film ? film[(size_t)b * 2 * C + c] + 1.f : 1.f;
        m1 += gamma[c] * sc * csum[((size_t)b * C + c) * 3 + 0];
        m2 += gamma[c] * sc * csum[((size_t)b * C + c) * 3 + 1];
    }
    m1 /= cnt;
    m2 /= cnt;
    // u = x*A + Bc ; dx = dout*silu'(u)*Gs - K0 - x*K1  with Gs = rstd*gamma*(sc+1), K1 = rstd^2*m2,
    // K0 = rstd*m1 - mean*K1
    float A[8], Bc[8], Gs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = slot * 8 + i;
        float ga = gamma[c] * rstd, be = beta[c] - mean * rstd * gamma[c];
        float sc = 1.f;
        if (film) {
            sc = film[(size_t)b * 2 * C + c] + 1.f;
            const float sh = film[(size_t)b * 2 * C + C + c];
            ga *= sc;
            be = be * sc + sh;
        }
        A[i] = 0.5f * ga;  // u / 2 (see dsilu_R_from_half)
        Bc[i] = 0.5f * be;
        Gs[i] = 0.5f * rstd * gamma[c] * sc;
    }
    const float K1 = rstd * rstd * m2, K0 = rstd * m1 - mean * K1;
    const size_t base = (size_t)b * P * C + slot * 8;
    const long long stride = (long long)gridDim.x * pix_per_iter;
    uint4 nu[PF], nd[PF];
    auto issue = [&](long long p) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const long long pk = p + k * stride;
            if (pk < P) {
                nu[k] = ld_stream16(x + base + pk * C);
                nd[k] = ld_stream16(dout + base + pk * C);
            }
        }
    };
    long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix;
    if (p < P) issue(p);
    // parameter gradients (the formulas of gn_bwd_params_kernel) ACCUMULATED by the first block of each
    // sample: saves a launch per GroupNorm when the caller hands in the parameters' .grad buffers
    if (dgamma != nullptr && blockIdx.x == 0 && (int)threadIdx.x < vec_per_pix) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = slot * 8 + i;
            const float* s3 = csum + ((size_t)b * C + c) * 3;
            const float sc = film ? film[(size_t)b * 2 * C + c] + 1.f : 1.f;
            atomicAdd(dbeta + c, sc * s3[0]);
            atomicAdd(dgamma + c, sc * s3[1]);
            if (dfilm) {
                dfilm[(size_t)b * 2 * C + c] = gamma[c] * s3[1] + beta[c] * s3[0];
                dfilm[(size_t)b * 2 * C + C + c] = s3[0];
            }
            if (dcbias)
                atomicAdd(dcbias + c, rstd * (gamma[c] * sc * s3[0] - (float)P * m1) -
                                          rstd * rstd * m2 * (s3[2] - (float)P * mean));
        }
    }
    for (; p < P; p += PF * stride) {
        uint4 u[PF], d[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            u[k] = nu[k];
            d[k] = nd[k];
        }
        if (p + PF * stride < P) issue(p + PF * stride);
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const long long pk = p + k * stride;
            if (pk >= P) break;
            const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
            const uint32_t dw[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h2(w[i]), dd = unpack_h2(dw[i]);
                // dout silu'(u) Gs = e (1 + R) with e = dout * Gs / 2 (Gs already halved)
                const float e0 = dd.x * Gs[2 * i], e1 = dd.y * Gs[2 * i + 1];
                const float o0 = fmaf(e0, dsilu_R_from_half(fmaf(f.x, A[2 * i], Bc[2 * i])), e0) + fmaf(-K1, f.x, -K0);
                const float o1 = fmaf(e1, dsilu_R_from_half(fmaf(f.y, A[2 * i + 1], Bc[2 * i + 1])), e1) + fmaf(-K1, f.y, -K0);
                o[i] = pack_h2(o0, o1);
            }
            *reinterpret_cast<uint4*>(dx + base + pk * C) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// Parameter gradients from the per-(b,c) sums T0 = sum du, T1 = sum du*xhat, T2 = sum x:
//   dgamma[c] = sum_b (sc+1) T1,  dbeta[c] = sum_b (sc+1) T0,
//   dfilm[b][c] (scale) = gamma T1 + beta T0,  dfilm[b][C+c] (shift) = T0,
//   dcbias[c] = sum_b sum_p dx[b][p][c]  (the bias gradient of the convolution that produced x):
//       sum_p dx = rstd*(gamma (sc+1) T0 - P m1) - rstd^2 m2 (T2 - P mean)
__global__ void gn_bwd_params_kernel(const float* __restrict__ csum, const float* __restrict__ sums,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ film, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ dfilm,
                                     float* __restrict__ dcbias, int B, long long P, int C, int G, float eps,
                                     int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int cpg = C / G, g = c / cpg;
    const float cnt = (float)P * (float)cpg;
    float dg = 0.f, db = 0.f, dcb = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* s = csum + ((size_t)b * C + c) * 3;
        const float sc = film ? film[(size_t)b * 2 * C + c] + 1.f : 1.f;
        db += sc * s[0];
        dg += sc * s[1];
        if (dfilm) {
            dfilm[(size_t)b * 2 * C + c] = gamma[c] * s[1] + beta[c] * s[0];
            dfilm[(size_t)b * 2 * C + C + c] = s[0];
        }
        if (dcbias) {
            const float mean = sums[((size_t)b * G + g) * 2] / cnt;
            const float var = fmaxf(sums[((size_t)b * G + g) * 2 + 1] / cnt - mean * mean, 0.f);
            const float rstd = rsqrtf(var + eps);
            float m1 = 0.f, m2 = 0.f;
            for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
                const float scc = film ? film[(size_t)b * 2 * C + cc] + 1.f : 1.f;
                m1 += gamma[cc] * scc * csum[((size_t)b * C + cc) * 3 + 0];
                m2 += gamma[cc] * scc * csum[((size_t)b * C + cc) * 3 + 1];
            }
            m1 /= cnt;
            m2 /= cnt;
            dcb += rstd * (gamma[c] * sc * s[0] - (float)P * m1) - rstd * rstd * m2 * (s[2] - (float)P * mean);
        }
    }
    if (accumulate) {
        dgamma[c] += dg;
        dbeta[c] += db;
        if (dcbias) dcbias[c] += dcb;
    } else {
        dgamma[c] = dg;
        dbeta[c] = db;
        if (dcbias) dcbias[c] = dcb;
    }
}

// ------------------------------------------------------------------------------------------------
// channel LayerNorm (gain only).  LPR lanes share one pixel row (LPR = min(32, C/8)), so a warp
// normalises 32/LPR rows at once and no lane idles at C = 64 / 128; lanes own 8 channels per
// 8*LPR-channel slab (NV slabs for C > 256).
// ------------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float sub_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int LPR, int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const h16* __restrict__ x, const float* __restrict__ gamma, h16* __restrict__ out,
              long long M, int C, float eps) {
    pdl_trigger();
    pdl_wait();
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sl = lane % LPR;
    const long long rows_per_blk = 8 * RPW;
    float gm[NV][8];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) gm[j][i] = gamma[j * 8 * LPR + sl * 8 + i];
    const float invC = 1.f / C;
    for (long long row0 = (long long)blockIdx.x * rows_per_blk; row0 < M; row0 += (long long)gridDim.x * rows_per_blk) {
        const long long row = row0 + (threadIdx.x >> 5) * RPW + lane / LPR;
        const bool valid = row < M;
        const long long rr = valid ? row : 0;
        Vec8 v[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            v[j] = load8(x + rr * C + j * 8 * LPR + sl * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[j].v[i];
        }
        const float mean = sub_sum<LPR>(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = v[j].v[i] - mean;
                q = fmaf(d, d, q);
            }
        const float rstd = rsqrtf(sub_sum<LPR>(q) * invC + eps);
        if (valid) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                Vec8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) o.v[i] = (v[j].v[i] - mean) * rstd * gm[j][i];
                store8(out + row * C + j * 8 * LPR + sl * 8, o);
            }
        }
    }
}

// dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) (+ dres), g = dy*gamma; dgamma[c] += sum_rows dy*xhat
template <int LPR, int NV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const h16* __restrict__ x, const float* __restrict__ gamma,
              const h16* __restrict__ dy, const h16* __restrict__ dres,
              h16* __restrict__ dx, float* __restrict__ dgamma, long long M, int C, float eps) {
    pdl_trigger();
    pdl_wait();
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sl = lane % LPR;
    const long long rows_per_blk = 8 * RPW;
    float gm[NV][8], dg[NV][8];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            gm[j][i] = gamma[j * 8 * LPR + sl * 8 + i];
            dg[j][i] = 0.f;
        }
    const float invC = 1.f / C;
    for (long long row0 = (long long)blockIdx.x * rows_per_blk; row0 < M; row0 += (long long)gridDim.x * rows_per_blk) {
        const long long row = row0 + (threadIdx.x >> 5) * RPW + lane / LPR;
        const bool valid = row < M;
        const long long rr = valid ? row : 0;
        Vec8 v[NV], d[NV], r[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const size_t off = rr * C + j * 8 * LPR + sl * 8;
            v[j] = load8(x + off);
            d[j] = load8(dy + off);
            if (dres) r[j] = load8(dres + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[j].v[i];
        }
        const float mean = sub_sum<LPR>(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float t = v[j].v[i] - mean;
                q = fmaf(t, t, q);
            }
        const float rstd = rsqrtf(sub_sum<LPR>(q) * invC + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (v[j].v[i] - mean) * rstd;
                const float g = d[j].v[i] * gm[j][i];
                if (valid) dg[j][i] = fmaf(d[j].v[i], xh, dg[j][i]);
                v[j].v[i] = xh;
                sg += g;
                sgx = fmaf(g, xh, sgx);
            }
        const float mg = sub_sum<LPR>(sg) * invC, mgx = sub_sum<LPR>(sgx) * invC;
        if (valid) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                Vec8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    o.v[i] = rstd * (d[j].v[i] * gm[j][i] - mg - v[j].v[i] * mgx);
                    if (dres) o.v[i] += r[j].v[i];
                }
                store8(dx + row * C + j * 8 * LPR + sl * 8, o);
            }
        }
    }
    // block-level reduction of dgamma: threads with equal (lane % LPR) own the same channels
    __shared__ float sh[256][NV * 8 + 1];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[threadIdx.x][j * 8 + i] = dg[j][i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        const int j = c / (8 * LPR), l = (c >> 3) % LPR, i = c & 7;
        float s = 0.f;
        for (int t = l; t < 256; t += LPR) s += sh[t][j * 8 + i];
        atomicAdd(&dgamma[c], s);
    }
}

// Streaming kernels want every SM busy for the whole launch: a grid of ~4 resident blocks per SM in
// total (over all `batches` of grid.y), each thread looping many trips, instead of one trip per block
// (whose latency would be the whole kernel) -- no wave quantisation, no tail.
// `resident` = blocks of this kernel that fit on the whole GPU at once (SMs x occupancy): the grid is
// exactly one resident wave, so there is no partially filled second wave.
static int norm_grid(long long work_items, int per_block, int batches = 1, int resident = 148 * 4) {
    long long blocks = (work_items + per_block - 1) / per_block;
    long long cap = resident / (batches < 1 ? 1 : batches);
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
template <typename Kern>
static int resident_blocks(Kern kern, int threads, int dyn_smem = 0) {
    int per_sm = 0, dev = 0, sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 2;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * per_sm;
}

}  // namespace cesm

using namespace cesm;

#define GN_CHECK(C, G)                                                                                   \
    CESM_REQUIRE((C) >= 64 && (C) <= 2048 && 2048 % (C) == 0 && (G) >= 1 && (G) <= 256 && (C) % (G) == 0 && \
                     ((C) / (G)) % 8 == 0,                                                               \
                 "GroupNorm needs C in {64..2048, divides 2048} and (C/G) %% 8 == 0 (C=%d G=%d)", (C), (G))

extern "C" int cesm_gn_stats(const void* x, float* sums, int B, long long P, int C, int G, void* stream) {
    GN_CHECK(C, G);
    CESM_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * B * G, as_stream(stream)));
    const int per_block = kNormThreads / (C / 8) * UNR;
    static const int res = resident_blocks(gn_stats_kernel, kNormThreads);
    dim3 grid(norm_grid(P, per_block, B, res), B);
    launch_pdl(gn_stats_kernel, grid, kNormThreads, 0, as_stream(stream), (const h16*)x, sums, P, C, G);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_gn_apply_fwd(const void* x, const float* sums, const float* gamma, const float* beta,
                                 const float* film, const void* residual, void* out, int B, long long P, int C,
                                 int G, float eps, void* stream) {
    GN_CHECK(C, G);
    const int per_block = kNormThreads / (C / 8) * UNR;
    static const int res = resident_blocks(gn_apply_fwd_kernel, kNormThreads);
    dim3 grid(norm_grid(P, per_block, B, res), B);
    launch_pdl(gn_apply_fwd_kernel, grid, kNormThreads, 0, as_stream(stream), 
        (const h16*)x, sums, gamma, beta, film, (const h16*)residual, (h16*)out, P, C, G,
        eps);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_gn_bwd(const void* x, const void* dout, const float* sums, const float* gamma, const float* beta,
                           const float* film, float* csum /* [B][C][3] scratch */, void* dx, float* dgamma,
                           float* dbeta, float* dfilm /* [B][2C] or NULL */, float* dconv_bias /* [C] or NULL */,
                           int B, long long P, int C, int G, float eps, int accumulate_params, void* stream) {
    GN_CHECK(C, G);
    cudaStream_t st = as_stream(stream);
    CESM_ZERO_SCRATCH(csum, sizeof(float) * 3 * B * C, st);
    const int per_block = kNormThreads / (C / 8) * UNR;
    constexpr int kRing = GN_STAGES * 2 * PF * kNormThreads * 16;  // 48 KB (+ 9 KB static: needs the opt-in)
    static bool ring_cfg = false;
    if (!ring_cfg) {
        CESM_CHECK_CUDA(cudaFuncSetAttribute(gn_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRing));
        ring_cfg = true;
    }
    static const int res_r = resident_blocks(gn_bwd_reduce_kernel, kNormThreads, kRing);
    static const int res_a = resident_blocks(gn_bwd_apply_kernel, kNormThreads);
    dim3 grid(norm_grid(P, per_block, B, res_r), B);
    dim3 grid_a(norm_grid(P, per_block, B, res_a), B);
    launch_pdl(gn_bwd_reduce_kernel, grid, kNormThreads, kRing, st, (const h16*)x, (const h16*)dout, sums,
                                                        gamma, beta, film, csum, P, C, G, eps);
    CESM_CHECK_LAUNCH();
    const bool fused = accumulate_params != 0;  // accumulate mode: the apply kernel adds the parameter gradients
    launch_pdl(gn_bwd_apply_kernel, grid_a, kNormThreads, 0, st, (const h16*)x, (const h16*)dout, sums, gamma,
                                                       beta, film, csum, (h16*)dx, P, C, G, eps,
                                                       fused ? dgamma : nullptr, dbeta, dfilm, dconv_bias);
    CESM_CHECK_LAUNCH();
    if (!fused) {
        launch_pdl(gn_bwd_params_kernel, ceil_div(C, 128), 128, 0, st, csum, sums, gamma, beta, film, dgamma, dbeta, dfilm,
                                                               dconv_bias, B, P, C, G, eps, 0);
        CESM_CHECK_LAUNCH();
    }
    return CESM_OK;
}

template <int LPR, int NV>
static void ln_launch_fwd(int /*grid*/, cudaStream_t st, const void* x, const float* gamma, void* out, long long M,
                          int C, float eps) {
    static const int res = resident_blocks(ln_fwd_kernel<LPR, NV>, 256);
    const int grid = norm_grid(M, 8 * (32 / LPR), 1, res);
    launch_pdl(ln_fwd_kernel<LPR, NV>, grid, 256, 0, st, (const h16*)x, gamma, (h16*)out, M, C, eps);
}
template <int LPR, int NV>
static void ln_launch_bwd(int /*grid*/, cudaStream_t st, const void* x, const float* gamma, const void* dy,
                          const void* dres, void* dx, float* dgamma, long long M, int C, float eps) {
    static const int res = resident_blocks(ln_bwd_kernel<LPR, NV>, 256);
    const int grid = norm_grid(M, 8 * (32 / LPR), 1, res);
    launch_pdl(ln_bwd_kernel<LPR, NV>, grid, 256, 0, st, (const h16*)x, gamma, (const h16*)dy,
                                                 (const h16*)dres, (h16*)dx, dgamma, M, C, eps);
}
// C in {64, 128, 256, 512, 1024}: LPR = min(32, C/8), NV = C / (8*LPR)
#define LN_DISPATCH(FN, ...)                                   \
    do {                                                       \
        switch (C) {                                           \
            case 64: FN<8, 1>(__VA_ARGS__); break;             \
            case 128: FN<16, 1>(__VA_ARGS__); break;           \
            case 256: FN<32, 1>(__VA_ARGS__); break;           \
            case 512: FN<32, 2>(__VA_ARGS__); break;           \
            default: FN<32, 4>(__VA_ARGS__); break;            \
        }                                                      \
    } while (0)
#define LN_CHECK(C) \
    CESM_REQUIRE((C) == 64 || (C) == 128 || (C) == 256 || (C) == 512 || (C) == 1024, \
                 "LayerNorm supports C in {64,128,256,512,1024} (C=%d)", (C))

extern "C" int cesm_ln_fwd(const void* x, const float* gamma, void* out, long long M, int C, float eps, void* stream) {
    LN_CHECK(C);
    cudaStream_t st = as_stream(stream);
    const int lpr = C / 8 < 32 ? C / 8 : 32;
    const int grid = norm_grid(M, 8 * (32 / lpr) * 2);
    LN_DISPATCH(ln_launch_fwd, grid, st, x, gamma, out, M, C, eps);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_ln_bwd(const void* x, const float* gamma, const void* dy, const void* dres, void* dx,
                           float* dgamma, long long M, int C, float eps, int accumulate, void* stream) {
    LN_CHECK(C);
    cudaStream_t st = as_stream(stream);
    if (!accumulate) CESM_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st));
    const int lpr = C / 8 < 32 ? C / 8 : 32;
    const int grid = norm_grid(M, 8 * (32 / lpr) * 4);
    LN_DISPATCH(ln_launch_bwd, grid, st, x, gamma, dy, dres, dx, dgamma, M, C, eps);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}
