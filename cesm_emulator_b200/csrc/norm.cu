// Bandwidth-bound normalisation kernels (bf16 channels-last activations, fp32 math/statistics).
//
//  GroupNorm + FiLM + SiLU (+ residual)      reference: video_net.py:216-227 (Block), :265 (+res)
//  channel LayerNorm (gain only)              reference: video_net.py:78-87
//
// Layout: x[b][p][c], p = (frame, row, col) flattened, c fastest.  Every thread owns one 16-byte
// vector (8 bf16 channels) of a pixel, so global accesses are fully coalesced 128-bit
// transactions; reductions go thread -> shared memory -> one fp32 atomic per (block, slot).
#include "api_common.h"
#include "common.cuh"

namespace cesm {

static constexpr int kNormThreads = 256;

struct Vec8 {
    float v[8];
};
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    Vec8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
    return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]);
    u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]);
    u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: sums[b][g] = (sum x, sum x^2) over the group's channels and all pixels.
// grid = (blocks_per_sample, B).  Requires (C/G) % 8 == 0 and 2048 % C == 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNormThreads)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums, long long P, int C, int G) {
    const int b = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int slot = threadIdx.x % vec_per_pix;       // fixed channel vector of this thread
    const int pix_per_iter = kNormThreads / vec_per_pix;
    const __nv_bfloat16* xb = x + (size_t)b * P * C;
    float s = 0.f, ss = 0.f;
    for (long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix; p < P;
         p += (long long)gridDim.x * pix_per_iter) {
        Vec8 v = load8(xb + p * C + slot * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += v.v[i];
            ss += v.v[i] * v.v[i];
        }
    }
    __shared__ float sh[2][kNormThreads];
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = ss;
    __syncthreads();
    // one thread per group gathers every slot that maps to it
    if (threadIdx.x < G) {
        const int cpg8 = (C / G) >> 3;  // channel vectors per group
        float a = 0.f, q = 0.f;
        for (int t = 0; t < kNormThreads; ++t) {
            if ((t % vec_per_pix) / cpg8 == (int)threadIdx.x) {
                a += sh[0][t];
                q += sh[1][t];
            }
        }
        atomicAdd(&sums[((size_t)b * G + threadIdx.x) * 2 + 0], a);
        atomicAdd(&sums[((size_t)b * G + threadIdx.x) * 2 + 1], q);
    }
}

// out = silu(((x-mean)*rstd*gamma + beta) * (scale+1) + shift) (+ residual)
__global__ void __launch_bounds__(kNormThreads)
gn_apply_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ sums,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ film,  // [B][2C] (scale | shift) or null
                    const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out, long long P, int C,
                    int G, float eps) {
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
    const size_t base = (size_t)b * P * C;
    for (long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix; p < P;
         p += (long long)gridDim.x * pix_per_iter) {
        const size_t off = base + p * C + slot * 8;
        Vec8 v = load8(x + off);
        Vec8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = silu_f(v.v[i] * A[i] + Bc[i]);
        if (residual) {
            Vec8 r = load8(residual + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] += r.v[i];
        }
        store8(out + off, o);
    }
}

// Backward pass 1: per (b, c) sums over pixels of
//   [0] dz, [1] dz*xhat, [2] du, [3] du*z    with z = xhat*gamma+beta, u = z*(sc+1)+sh, du = dout*silu'(u),
//   dz = du*(sc+1).
__global__ void __launch_bounds__(kNormThreads)
gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ sums, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ film, float* __restrict__ csum,
                     long long P, int C, int G, float eps) {
    const int b = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int slot = threadIdx.x % vec_per_pix;
    const int pix_per_iter = kNormThreads / vec_per_pix;
    const int g = (slot * 8) / (C / G);
    const float cnt = (float)P * (float)(C / G);
    const float mean = sums[((size_t)b * G + g) * 2] / cnt;
    const float var = fmaxf(sums[((size_t)b * G + g) * 2 + 1] / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float ga[8], be[8], sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = slot * 8 + i;
        ga[i] = gamma[c];
        be[i] = beta[c];
        sc[i] = film ? film[(size_t)b * 2 * C + c] + 1.f : 1.f;
        sh[i] = film ? film[(size_t)b * 2 * C + C + c] : 0.f;
    }
    float acc[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    const size_t base = (size_t)b * P * C;
    for (long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix; p < P;
         p += (long long)gridDim.x * pix_per_iter) {
        const size_t off = base + p * C + slot * 8;
        Vec8 v = load8(x + off);
        Vec8 d = load8(dout + off);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xh = (v.v[i] - mean) * rstd;
            const float z = xh * ga[i] + be[i];
            const float u = z * sc[i] + sh[i];
            const float du = d.v[i] * dsilu_f(u);
            const float dz = du * sc[i];
            acc[0][i] += dz;
            acc[1][i] += dz * xh;
            acc[2][i] += du;
            acc[3][i] += du * z;
        }
    }
    // reduce threads that share a channel slot (stride vec_per_pix) through shared memory
    __shared__ float red[kNormThreads][9];
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[k][i];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += kNormThreads) {  // thread t sums channels t, t+256, ...
            float s = 0.f;
            for (int t = c >> 3; t < kNormThreads; t += vec_per_pix) s += red[t][c & 7];
            atomicAdd(&csum[((size_t)b * C + c) * 4 + k], s);
        }
        __syncthreads();
    }
}

// Backward pass 2: dx = rstd * (gamma*dz - m1 - xhat*m2), m1/m2 = group means of gamma*dz and gamma*dz*xhat.
__global__ void __launch_bounds__(kNormThreads)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dout,
                    const float* __restrict__ sums, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ film,
                    const float* __restrict__ csum, __nv_bfloat16* __restrict__ dx, long long P, int C, int G,
                    float eps) {
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
        m1 += gamma[c] * csum[((size_t)b * C + c) * 4 + 0];
        m2 += gamma[c] * csum[((size_t)b * C + c) * 4 + 1];
    }
    m1 /= cnt;
    m2 /= cnt;
    float ga[8], be[8], sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = slot * 8 + i;
        ga[i] = gamma[c];
        be[i] = beta[c];
        sc[i] = film ? film[(size_t)b * 2 * C + c] + 1.f : 1.f;
        sh[i] = film ? film[(size_t)b * 2 * C + C + c] : 0.f;
    }
    const size_t base = (size_t)b * P * C;
    for (long long p = (long long)blockIdx.x * pix_per_iter + threadIdx.x / vec_per_pix; p < P;
         p += (long long)gridDim.x * pix_per_iter) {
        const size_t off = base + p * C + slot * 8;
        Vec8 v = load8(x + off);
        Vec8 d = load8(dout + off);
        Vec8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xh = (v.v[i] - mean) * rstd;
            const float u = (xh * ga[i] + be[i]) * sc[i] + sh[i];
            const float dz = d.v[i] * dsilu_f(u) * sc[i];
            o.v[i] = rstd * (ga[i] * dz - m1 - xh * m2);
        }
        store8(dx + off, o);
    }
}

// Parameter gradients from the per-(b,c) sums: dgamma[c] = sum_b S1, dbeta[c] = sum_b S0,
// dfilm[b][c] = S3 (scale), dfilm[b][C+c] = S2 (shift).
__global__ void gn_bwd_params_kernel(const float* __restrict__ csum, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ dfilm, int B, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float dg = 0.f, db = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* s = csum + ((size_t)b * C + c) * 4;
        db += s[0];
        dg += s[1];
        if (dfilm) {
            dfilm[(size_t)b * 2 * C + c] = s[3];
            dfilm[(size_t)b * 2 * C + C + c] = s[2];
        }
    }
    dgamma[c] = dg;
    dbeta[c] = db;
}

// ------------------------------------------------------------------------------------------------
// channel LayerNorm: one warp per pixel row of C channels (C <= 512, C % 64 == 0)
// ------------------------------------------------------------------------------------------------
// Each lane owns channels {lane*8 .. lane*8+7} + 256*j for j < NV (NV = ceil(C/256)); lanes whose
// vector starts beyond C are idle for that j.
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ out,
              long long M, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    for (long long row = row0; row < M; row += (long long)gridDim.x * 8) {
        Vec8 v[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
                v[j] = load8(x + row * C + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) s += v[j].v[i];
            }
        }
        const float mean = warp_sum(s) / C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = v[j].v[i] - mean;
                    q += d * d;
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
                Vec8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) o.v[i] = (v[j].v[i] - mean) * rstd * gamma[c + i];
                store8(out + row * C + c, o);
            }
        }
    }
}

// dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) (+ dres), g = dy*gamma; dgamma[c] += sum_rows dy*xhat
template <int NV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dres,
              __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, long long M, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float dg[NV][8];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) dg[j][i] = 0.f;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < M; row += (long long)gridDim.x * 8) {
        Vec8 v[NV], d[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
                v[j] = load8(x + row * C + c);
                d[j] = load8(dy + row * C + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) s += v[j].v[i];
            }
        }
        const float mean = warp_sum(s) / C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float t = v[j].v[i] - mean;
                    q += t * t;
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / C + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xh = (v[j].v[i] - mean) * rstd;
                    const float g = d[j].v[i] * gamma[c + i];
                    dg[j][i] += d[j].v[i] * xh;
                    sg += g;
                    sgx += g * xh;
                }
            }
        }
        const float mg = warp_sum(sg) / C, mgx = warp_sum(sgx) / C;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = j * 256 + lane * 8;
            if (c < C) {
                Vec8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xh = (v[j].v[i] - mean) * rstd;
                    o.v[i] = rstd * (d[j].v[i] * gamma[c + i] - mg - xh * mgx);
                }
                if (dres) {
                    Vec8 r = load8(dres + row * C + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o.v[i] += r.v[i];
                }
                store8(dx + row * C + c, o);
            }
        }
    }
    // block-level reduction of dgamma: 8 warps -> shared -> one atomic per channel per block
    __shared__ float sh[8][NV * 256];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[warp][j * 256 + lane * 8 + i] = dg[j][i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[w][c];
        atomicAdd(&dgamma[c], s);
    }
}

static int norm_grid(long long work_items, int per_block) {
    long long blocks = (work_items + per_block - 1) / per_block;
    const long long cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
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
    const int per_block = kNormThreads / (C / 8) * 8;
    dim3 grid(norm_grid(P, per_block), B);
    gn_stats_kernel<<<grid, kNormThreads, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, sums, P, C, G);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_gn_apply_fwd(const void* x, const float* sums, const float* gamma, const float* beta,
                                 const float* film, const void* residual, void* out, int B, long long P, int C,
                                 int G, float eps, void* stream) {
    GN_CHECK(C, G);
    const int per_block = kNormThreads / (C / 8) * 4;
    dim3 grid(norm_grid(P, per_block), B);
    gn_apply_fwd_kernel<<<grid, kNormThreads, 0, as_stream(stream)>>>(
        (const __nv_bfloat16*)x, sums, gamma, beta, film, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, P, C, G,
        eps);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_gn_bwd(const void* x, const void* dout, const float* sums, const float* gamma, const float* beta,
                           const float* film, float* csum /* [B][C][4] scratch */, void* dx, float* dgamma,
                           float* dbeta, float* dfilm /* [B][2C] or NULL */, int B, long long P, int C, int G,
                           float eps, void* stream) {
    GN_CHECK(C, G);
    cudaStream_t st = as_stream(stream);
    CESM_CHECK_CUDA(cudaMemsetAsync(csum, 0, sizeof(float) * 4 * B * C, st));
    const int per_block = kNormThreads / (C / 8) * 8;
    dim3 grid(norm_grid(P, per_block), B);
    gn_bwd_reduce_kernel<<<grid, kNormThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dout, sums,
                                                        gamma, beta, film, csum, P, C, G, eps);
    CESM_CHECK_LAUNCH();
    gn_bwd_apply_kernel<<<grid, kNormThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dout, sums, gamma,
                                                       beta, film, csum, (__nv_bfloat16*)dx, P, C, G, eps);
    CESM_CHECK_LAUNCH();
    gn_bwd_params_kernel<<<ceil_div(C, 128), 128, 0, st>>>(csum, dgamma, dbeta, dfilm, B, C);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

#define LN_DISPATCH(KERNEL, ...)                                                  \
    do {                                                                          \
        const int nv = ceil_div(C, 256);                                          \
        if (nv == 1) KERNEL<1><<<grid, 256, 0, st>>>(__VA_ARGS__);                \
        else if (nv == 2) KERNEL<2><<<grid, 256, 0, st>>>(__VA_ARGS__);           \
        else KERNEL<4><<<grid, 256, 0, st>>>(__VA_ARGS__);                        \
    } while (0)

extern "C" int cesm_ln_fwd(const void* x, const float* gamma, void* out, long long M, int C, float eps, void* stream) {
    CESM_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1024, "LayerNorm needs C %% 8 == 0 and C <= 1024 (C=%d)", C);
    cudaStream_t st = as_stream(stream);
    const int grid = norm_grid(M, 8 * 4);
    LN_DISPATCH(ln_fwd_kernel, (const __nv_bfloat16*)x, gamma, (__nv_bfloat16*)out, M, C, eps);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}

extern "C" int cesm_ln_bwd(const void* x, const float* gamma, const void* dy, const void* dres, void* dx,
                           float* dgamma, long long M, int C, float eps, void* stream) {
    CESM_REQUIRE(C % 8 == 0 && C >= 8 && C <= 1024, "LayerNorm needs C %% 8 == 0 and C <= 1024 (C=%d)", C);
    cudaStream_t st = as_stream(stream);
    CESM_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st));
    const int grid = norm_grid(M, 8 * 8);
    LN_DISPATCH(ln_bwd_kernel, (const __nv_bfloat16*)x, gamma, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dres,
                (__nv_bfloat16*)dx, dgamma, M, C, eps);
    CESM_CHECK_LAUNCH();
    return CESM_OK;
}
