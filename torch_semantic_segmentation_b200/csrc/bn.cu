// BatchNorm (training statistics / apply / backward) and the small NHWC elementwise
// kernels around it.  All of them are pure HBM streaming: 8 channels per thread
// (128-bit bf16 accesses), grid-stride over (row, channel-group) items, rows may be
// pitched (ld*) so that slices of a concat buffer can be read and written in place.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int stream_grid(int64_t items, int per_sm = 8) {
    int64_t want = ceil_div64(items, kThreads);
    int64_t cap = (int64_t)tss_num_sms() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// launch shape of the (channel group, row lane) kernels: CG = C/8 groups x PL row lanes per CTA, U rows in flight per thread
struct RowsLaunch {
    int threads, PL;
    int64_t M;
    int grid(int U, int per_sm = 8) const {
        int64_t want = ceil_div64(M, (int64_t)PL * U);          // >= 1 batch of U rows per row lane
        const int64_t cap = (int64_t)tss_num_sms() * per_sm;
        if (want < 1) want = 1;
        return (int)(want < cap ? want : cap);
    }
};
inline RowsLaunch rows_launch(int64_t M, int C) {
    RowsLaunch r;
    const int CG = C / 8;
    r.M = M;
    r.PL = CG <= kThreads ? kThreads / CG : 0;
    r.threads = r.PL * CG;
    return r;
}

__device__ __forceinline__ void ldg8f(const float* __restrict__ p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ------------------------------------------------------------------ finalize / fold --
__global__ void bn_finalize_kernel(double* __restrict__ stats, double inv_count, double unbias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   int64_t* __restrict__ nbt, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int C,
                                   int clear_n) {
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt != nullptr) *nbt += 1;
    if (c >= C) return;
    const double st1 = stats[c], st2 = stats[C + c];
    // "consume and clear": thread c only ever touches indices == c (mod C), so nobody clears what
    // another thread still has to read; the scratch is zero again for its next use
    for (int i = c; i < clear_n; i += C) stats[i] = 0.0;
    const double mean = st1 * inv_count;
    double var = st2 * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma != nullptr ? gamma[c] : 1.f;
    const float b = beta != nullptr ? beta[c] : 0.f;
    const float sc = g * rstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    mean_out[c] = (float)mean;
    rstd_out[c] = rstd;
    if (running_mean != nullptr) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
    }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = (gamma != nullptr ? gamma[c] : 1.f) / sqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = (beta != nullptr ? beta[c] : 0.f) - rm[c] * sc;
}

// ------------------------------------------------------------------ apply -------------
// Thread = (8-channel group, row lane): the per-channel constants are loaded ONCE into registers and the thread walks
// rows with U independent 128-bit loads in flight per operand.  (Round 1 mapped a flat item index to (row, group) with a
// 64-bit division and re-loaded scale / shift for every 16 bytes of data: 4 parameter loads per data load, and the
// kernels sat at 0.36 of the HBM roofline with 10-16 us floors on tensors that fit in L2.)
template <typename T, int U, bool kExtra>       // kExtra: a second normalised branch (y2) and / or a residual are added
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                const T* __restrict__ y2, const float* __restrict__ scale2, const float* __restrict__ shift2,
                const T* __restrict__ res, T* __restrict__ z, int64_t M, int C,
                int64_t ldy, int64_t ldy2, int64_t ldr, int64_t ldz, int relu, int PL) {
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG;
    const int pl = threadIdx.x / CG;
    const int c0 = cg * 8;
    pdl_wait();
    if (pl >= PL) return;
    float sc[8], sh[8], sc2[8], sh2[8];
    ldg8f(scale + c0, sc);
    ldg8f(shift + c0, sh);
    if (kExtra && y2 != nullptr) { ldg8f(scale2 + c0, sc2); ldg8f(shift2 + c0, sh2); }
    const int64_t step = (int64_t)gridDim.x * PL;
    for (int64_t m0 = (int64_t)blockIdx.x * PL + pl; m0 < M; m0 += U * step) {
        Raw8<T> ry[U], ry2[kExtra ? U : 1], rr[kExtra ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m < M) {
                ry[u].ld(y + m * ldy + c0);
                if (kExtra && y2 != nullptr) ry2[u].ld(y2 + m * ldy2 + c0);
                if (kExtra && res != nullptr) rr[u].ld(res + m * ldr + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m >= M) break;
            float v[8];
            ry[u].get(v);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], sc[e], sh[e]);
            if (kExtra && y2 != nullptr) {
                float w[8];
                ry2[u].get(w);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += fmaf(w[e], sc2[e], sh2[e]);
            }
            if (kExtra && res != nullptr) {
                float w[8];
                rr[u].get(w);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += w[e];
            }
            if (relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            store8(z + m * ldz + c0, v);
        }
    }
}

// ------------------------------------------------------------------ backward ----------
// pass 1: per-channel sums of g and g*xhat.  Block = CG channel groups x PL row lanes, same thread mapping as the apply
// kernels.  (The round-1 kernel needed 121 registers -- 2 CTAs per SM -- and reduced inside the CTA with shared-memory
// float atomics, CAS loops on sm_100: 2.3-2.5 TB/s on the large maps, 8-12 us floors on the small ones.  Here rstd
// is applied once at the end, the mask mode is a template parameter, three CTAs per SM are resident and the lanes
// of a CTA are summed through a [PL][C] shared-memory array: no atomics in front of the one global RED per channel.)
constexpr int kMaskNone = 0, kMaskZ = 1, kMaskY = 2;    // no ReLU / mask from the activated tensor z / recomputed from y

template <typename T, int U, int kMask>
__device__ __forceinline__ void bwd_reduce_rows(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                                                int64_t M, int64_t lddz, int64_t ldz, int64_t ldy, int c0,
                                                int64_t first, int64_t step, const float (&mu)[8], const float (&sc)[8],
                                                const float (&sh)[8], float (&s1)[8], float (&s2)[8]) {
    for (int64_t m0 = first; m0 < M; m0 += U * step) {
        Raw8<T> rg[U], ry[U], rz[kMask == kMaskZ ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m < M) {
                rg[u].ld(dz + m * lddz + c0);
                ry[u].ld(y + m * ldy + c0);
                if (kMask == kMaskZ) rz[u].ld(z + m * ldz + c0);
            } else {
                rg[u].zero(); ry[u].zero();
                if (kMask == kMaskZ) rz[u].zero();
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float g[8], yy[8];
            rg[u].get(g);
            ry[u].get(yy);
            if (kMask == kMaskZ) {
                float zz[8];
                rz[u].get(zz);
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = zz[e] > 0.f ? g[e] : 0.f;
            } else if (kMask == kMaskY) {
                // the forward's own arithmetic (scale = gamma*rstd, shift = beta - mean*scale, fma): the identical mask
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = fmaf(yy[e], sc[e], sh[e]) > 0.f ? g[e] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s1[e] += g[e];
                s2[e] = fmaf(g[e], yy[e] - mu[e], s2[e]);        // rstd is applied to the finished sum
            }
        }
    }
}

// the CTA's lanes summed per channel through s_part[PL][C] (<= kThreads * 8 floats), then ONE global RED per channel
__device__ __forceinline__ void cta_channel_sums(const float (&v)[8], float* s_part, float* __restrict__ out, int C, int c0,
                                                 int pl, int PL) {
    float* mine = s_part + (size_t)pl * C + c0;
    *reinterpret_cast<float4*>(mine) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(mine + 4) = make_float4(v[4], v[5], v[6], v[7]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float a = 0.f;
        for (int p = 0; p < PL; ++p) a += s_part[(size_t)p * C + i];
        atomicAdd(out + i, a);
    }
    __syncthreads();                                             // s_part may be rewritten
}

template <int kMask>
__device__ __forceinline__ void bwd_constants(const float* __restrict__ mean, const float* __restrict__ rstd,
                                              const float* __restrict__ gamma, const float* __restrict__ beta, int c0,
                                              float (&mu)[8], float (&rs)[8], float (&sc)[8], float (&sh)[8]) {
    ldg8f(mean + c0, mu);
    ldg8f(rstd + c0, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sc[e] = (gamma != nullptr ? __ldg(gamma + c0 + e) : 1.f) * rs[e];
        sh[e] = (kMask == kMaskY && beta != nullptr ? __ldg(beta + c0 + e) : 0.f) - mu[e] * sc[e];
    }
}

template <typename T, int U, int kMask>
__global__ void __launch_bounds__(kThreads, 3)
bn_bwd_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ sums, int64_t M, int C, int64_t lddz, int64_t ldz, int64_t ldy, int PL) {
    __shared__ __align__(16) float s_part[kThreads * 8];
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG;
    const int pl = threadIdx.x / CG;
    const int c0 = cg * 8;
    pdl_wait();
    float s1[8], s2[8], mu[8], rs[8], sc[8], sh[8];
    zero8(s1); zero8(s2);
    bwd_constants<kMask>(mean, rstd, gamma, beta, c0, mu, rs, sc, sh);
    bwd_reduce_rows<T, U, kMask>(dz, z, y, M, lddz, ldz, ldy, c0, (int64_t)blockIdx.x * PL + pl, (int64_t)gridDim.x * PL,
                                 mu, sc, sh, s1, s2);
#pragma unroll
    for (int e = 0; e < 8; ++e) s2[e] *= rs[e];
    cta_channel_sums(s1, s_part, sums, C, c0, pl, PL);
    cta_channel_sums(s2, s_part, sums + C, C, c0, pl, PL);
}

// pass 2: dy = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)); optional dres = g.  Same thread mapping as bn_apply_kernel.
// Per-channel constants in registers: a = gamma*rstd (the forward's scale), k2 = rstd*sum(g*xhat)/n, c1 = sum(g)/n - mean*k2
// (and the forward's shift for the recomputed ReLU mask): dy = a * (g - c1 - y*k2).  Three CTAs per SM (<= 80 registers;
// the first version kept mean, k1, k2 and the shift apart: 112 registers, two CTAs, 3.4-4.1 TB/s on the 30-60 MB tensors).
template <typename T, int U, int kMask>         // kMaskZ: residual layers (mask from the activated tensor z, dres = g stored)
__global__ void __launch_bounds__(kThreads, 3)
bn_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ sums,
                    T* __restrict__ dy, T* __restrict__ dres, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int64_t M, int C, int64_t lddz, int64_t ldz, int64_t ldy,
                    int64_t lddy, int64_t lddres, float inv_m, int PL) {
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG;
    const int pl = threadIdx.x / CG;
    const int c0 = cg * 8;
    pdl_wait();
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta != nullptr) dbeta[c] += sums[c];
            if (dgamma != nullptr) dgamma[c] += sums[C + c];
        }
    }
    if (pl >= PL) return;
    constexpr bool mask_z = kMask == kMaskZ, mask_y = kMask == kMaskY;
    float a[8], c1[8], k2[8], sh[8];
    {
        float mu[8], rs[8], s1[8], s2[8];
        ldg8f(mean + c0, mu);
        ldg8f(rstd + c0, rs);
        ldg8f(sums + c0, s1);
        ldg8f(sums + C + c0, s2);
        if (gamma != nullptr) ldg8f(gamma + c0, a);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = 1.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            a[e] *= rs[e];                                   // = the forward's scale
            k2[e] = rs[e] * (s2[e] * inv_m);
            c1[e] = fmaf(-mu[e], k2[e], s1[e] * inv_m);
            sh[e] = mask_y ? (beta != nullptr ? __ldg(beta + c0 + e) : 0.f) - mu[e] * a[e] : 0.f;
        }
    }
    const int64_t step = (int64_t)gridDim.x * PL;
    for (int64_t m0 = (int64_t)blockIdx.x * PL + pl; m0 < M; m0 += U * step) {
        Raw8<T> rg[U], ry[U], rz[mask_z ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m < M) {
                rg[u].ld(dz + m * lddz + c0);
                ry[u].ld(y + m * ldy + c0);
                if (mask_z) rz[u].ld(z + m * ldz + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m >= M) break;
            float g[8], yy[8];
            rg[u].get(g);
            ry[u].get(yy);
            if (mask_z) {
                float zz[8];
                rz[u].get(zz);
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = zz[e] > 0.f ? g[e] : 0.f;
            } else if (mask_y) {
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = fmaf(yy[e], a[e], sh[e]) > 0.f ? g[e] : 0.f;
            }
            if (dres != nullptr) store8(dres + m * lddres + c0, g);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = a[e] * fmaf(-yy[e], k2[e], g[e] - c1[e]);
            store8(dy + m * lddy + c0, o);
        }
    }
}

// ------------------------------------------------------------------ backward in ONE launch ------
// Both passes of the BatchNorm backward in one kernel, for tensors that stay in L2 between them (dz and y together
// up to a few tens of MB: every layer at 1/16 and 1/32 resolution and the 64 / 128-channel maps at 1/8): pass 1 as
// above, a grid-wide barrier, pass 2 re-reads dz and y from L2 instead of HBM and the second launch (with its
// dependent-launch edge and its constant loads) disappears.  All CTAs must be resident at once: the launcher caps the
// grid at two CTAs per SM (<= 104 registers: room is left for a weight-gradient CTA beside them).  Whatever runs beside this kernel (weight gradients on the
// side stream) cannot depend on it, so it drains and every CTA gets its slot; the next kernel of this stream only
// becomes resident once every CTA here has passed griddepcontrol.launch_dependents, i.e. is resident itself.
//
// sync[0] = arrival counter (left at 0), sync[1] = generation (bumped by the last arrival), sync[2] = sticky error flag:
// a CTA that waits longer than ~1 s gives up, sets it, and from then on nobody waits (wrong sums instead of a hung GPU;
// ops.bn_backward checks the flag in debug mode, tests/test_fused_paths_gpu.py always).
__device__ __forceinline__ void grid_barrier(int* sync) {
    __syncthreads();
#ifndef TSS_HOST_EMU                                             // (the emulation runs one CTA: nothing to wait for)
    if (threadIdx.x == 0) {
        volatile int* vs = sync;
        const int gen = vs[1];                                   // read before arriving: it cannot change until this CTA has arrived too
        __threadfence();                                         // this CTA's REDs on the sums before its arrival
        if (atomicAdd(sync, 1) == (int)gridDim.x - 1) {
            vs[0] = 0;                                           // re-armed for the next launch
            __threadfence();
            atomicAdd(sync + 1, 1);
        } else {
            unsigned spins = 0;
            while (vs[1] == gen && vs[2] == 0) {
                __nanosleep(64);
                if (++spins > (1u << 21)) { atomicExch(sync + 2, 1); break; }
            }
        }
        __threadfence();
    }
    __syncthreads();
#endif
}

__device__ __forceinline__ void ldcg8f(const float* p, float (&v)[8]) {      // L2 (written by other SMs in this launch)
    float4 a = __ldcg(reinterpret_cast<const float4*>(p));
    float4 b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <typename T, int U, int kMask>         // kMaskZ: residual layers (mask from the activated tensor z, dres = g stored)
__global__ void __maxnreg__(96)         // launched with <= kThreads threads; two CTAs fit beside two weight-gradient CTAs (192 x 40 registers)
bn_bwd_onepass_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y,
                      const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* sums, T* __restrict__ dy, T* __restrict__ dres, float* __restrict__ dgamma,
                      float* __restrict__ dbeta, int64_t M, int C, int64_t lddz, int64_t ldz, int64_t ldy, int64_t lddy,
                      int64_t lddres, float inv_m, int PL, int* sync) {
    __shared__ __align__(16) float s_part[kThreads * 8];
    const int CG = C >> 3;
    const int cg = threadIdx.x % CG;
    const int pl = threadIdx.x / CG;
    const int c0 = cg * 8;
    pdl_wait();
    // live across pass 1: mu, a, sh, s1, s2; across pass 2: a, sh, c1, k2 (dy = a * (g - c1 - y * k2), c1 = k1 - mean * k2)
    float a[8], sh[8], c1[8], k2[8];
    const int64_t first = (int64_t)blockIdx.x * PL + pl, step = (int64_t)gridDim.x * PL;
    {
        float s1[8], s2[8], mu[8], rs[8];
        zero8(s1); zero8(s2);
        bwd_constants<kMask>(mean, rstd, gamma, beta, c0, mu, rs, a, sh);         // a = the forward's scale
        bwd_reduce_rows<T, U, kMask>(dz, z, y, M, lddz, ldz, ldy, c0, first, step, mu, a, sh, s1, s2);
        ldg8f(rstd + c0, rs);                                                      // (re-read: not kept across the loop)
#pragma unroll
        for (int e = 0; e < 8; ++e) s2[e] *= rs[e];
        cta_channel_sums(s1, s_part, sums, C, c0, pl, PL);
        cta_channel_sums(s2, s_part, sums + C, C, c0, pl, PL);
        grid_barrier(sync);
        ldcg8f(sums + c0, s1);
        ldcg8f(sums + C + c0, s2);
        ldg8f(mean + c0, mu);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            k2[e] = rs[e] * (s2[e] * inv_m);
            c1[e] = fmaf(-mu[e], k2[e], s1[e] * inv_m);
        }
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta != nullptr) dbeta[c] += __ldcg(sums + c);
            if (dgamma != nullptr) dgamma[c] += __ldcg(sums + C + c);
        }
    }
    // pass 2 walks this thread's rows backwards: what pass 1 touched last is what L2 still holds for certain
    if (first >= M) return;
    const int64_t batches = (M - first + U * step - 1) / (U * step);
    for (int64_t b = batches - 1; b >= 0; --b) {
        const int64_t m0 = first + b * U * step;
        Raw8<T> rg[U], ry[U], rz[kMask == kMaskZ ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m < M) {
                rg[u].ld(dz + m * lddz + c0);
                ry[u].ld(y + m * ldy + c0);
                if (kMask == kMaskZ) rz[u].ld(z + m * ldz + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t m = m0 + u * step;
            if (m >= M) break;
            float g[8], yy[8];
            rg[u].get(g);
            ry[u].get(yy);
            if (kMask == kMaskZ) {
                float zz[8];
                rz[u].get(zz);
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = zz[e] > 0.f ? g[e] : 0.f;
            } else if (kMask == kMaskY) {
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = fmaf(yy[e], a[e], sh[e]) > 0.f ? g[e] : 0.f;
            }
            if (dres != nullptr) store8(dres + m * lddres + c0, g);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = a[e] * fmaf(-yy[e], k2[e], g[e] - c1[e]);
            store8(dy + m * lddy + c0, o);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
relu_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ z, T* __restrict__ g, int64_t M, int C,
                int64_t lddz, int64_t ldz, int64_t ldg) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = M * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        const int64_t m = item / CG;
        const int c0 = (int)(item - m * CG) * 8;
        float a[8], b[8];
        load8(dz + m * lddz + c0, a);
        load8(z + m * ldz + c0, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = b[e] > 0.f ? a[e] : 0.f;
        store8(g + m * ldg + c0, a);
    }
}

// out = a + b (b may be nullptr: plain strided copy)
template <typename T>
__global__ void __launch_bounds__(kThreads)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t M, int C,
           int64_t lda, int64_t ldb, int64_t ldo) {
    pdl_wait();
    const int CG = C >> 3;
    const int64_t total = M * CG;
    for (int64_t item = (int64_t)blockIdx.x * kThreads + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * kThreads) {
        const int64_t m = item / CG;
        const int c0 = (int)(item - m * CG) * 8;
        float u[8];
        load8(a + m * lda + c0, u);
        if (b != nullptr) {
            float v[8];
            load8(b + m * ldb + c0, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) u[e] += v[e];
        }
        store8(out + m * ldo + c0, u);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n8) {
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kThreads) {
        float v[8];
        load8(src + i * 8, v);
        store8(dst + i * 8, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
scale_inplace_kernel(T* __restrict__ x, const float* __restrict__ s, int64_t n8) {
    pdl_wait();
    const float k = __ldg(s);
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kThreads) {
        float v[8];
        load8(x + i * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= k;
        store8(x + i * 8, v);
    }
}

int check_rows(const char* name, int64_t M, int C) {
    TSS_REQUIRE(M > 0, "%s: empty tensor (M=%lld)", name, (long long)M);
    TSS_REQUIRE(C > 0 && C % 8 == 0 && C <= 2048, "%s: C=%d must be a multiple of 8 (<= 2048)", name, C);
    return TSS_OK;
}

}  // namespace

extern "C" int tss_bn_finalize(double* stats, int64_t count, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, int64_t* num_batches_tracked,
                               float momentum, float eps, float* scale, float* shift, float* mean,
                               float* rstd, int C, int64_t clear_n, void* stream) {
    TSS_REQUIRE(C > 0, "bn_finalize: C=%d", C);
    TSS_REQUIRE(clear_n >= 0 && clear_n <= 64 * (int64_t)C, "bn_finalize: clear_n=%lld out of range", (long long)clear_n);
    // nn.BatchNorm2d raises "Expected more than 1 value per channel when training"
    TSS_REQUIRE(count > 1, "bn_finalize: Expected more than 1 value per channel when training, got %lld",
                (long long)count);
    TSS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats mismatch");
    const double unbias = (double)count / (double)(count - 1);
    tss_launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, 
        stats, 1.0 / (double)count, unbias, gamma, beta, running_mean, running_var, num_batches_tracked,
        momentum, eps, scale, shift, mean, rstd, C, (int)clear_n);
    TSS_LAUNCH_CHECK("bn_finalize");
    return TSS_OK;
}

extern "C" int tss_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float eps, float* scale, float* shift, int C,
                           void* stream) {
    TSS_REQUIRE(C > 0, "bn_fold: C=%d", C);
    tss_launch(bn_fold_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, gamma, beta, running_mean, running_var,
                                                                     eps, scale, shift, C);
    TSS_LAUNCH_CHECK("bn_fold");
    return TSS_OK;
}

extern "C" int tss_bn_apply(const void* y, const float* scale, const float* shift, const void* y2,
                            const float* scale2, const float* shift2, const void* res, void* z, int64_t M,
                            int C, int64_t ldy, int64_t ldy2, int64_t ldr, int64_t ldz, int flags, int dtype,
                            void* stream) {
    if (int e = check_rows("bn_apply", M, C)) return e;
    TSS_REQUIRE(y2 == nullptr || (scale2 != nullptr && shift2 != nullptr), "bn_apply: y2 without scale2/shift2");
    const RowsLaunch rl = rows_launch(M, C);
    TSS_REQUIRE(rl.threads > 0, "bn_apply: C=%d too large", C);
    TSS_DISPATCH_DTYPE(dtype, "bn_apply", {
        constexpr int U = sizeof(T) == 2 ? 4 : 2;
        auto kern = (y2 != nullptr || res != nullptr) ? bn_apply_kernel<T, U, true> : bn_apply_kernel<T, U, false>;
        tss_launch(kern, rl.grid(U), rl.threads, 0, (cudaStream_t)stream,
            (const T*)y, scale, shift, (const T*)y2, scale2, shift2, (const T*)res, (T*)z, M, C, ldy, ldy2,
            ldr, ldz, flags & TSS_EPI_RELU, rl.PL);
        TSS_LAUNCH_CHECK("bn_apply");
        return TSS_OK;
    });
}

extern "C" int tss_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* mean,
                                 const float* rstd, const float* gamma, const float* beta, float* sums,
                                 int64_t M, int C, int64_t lddz, int64_t ldz,
                                 int64_t ldy, int flags, int dtype, void* stream) {
    if (int e = check_rows("bn_bwd_reduce", M, C)) return e;
    const int relu = flags & TSS_EPI_RELU;
    const RowsLaunch rl = rows_launch(M, C);
    TSS_REQUIRE(rl.threads > 0, "bn_bwd_reduce: C=%d too large", C);
    const int mask = !relu ? kMaskNone : (z != nullptr ? kMaskZ : kMaskY);
    TSS_DISPATCH_DTYPE(dtype, "bn_bwd_reduce", {
        constexpr int U = sizeof(T) == 2 ? 4 : 2;
        constexpr int U3 = sizeof(T) == 2 ? 3 : 2;
        // (bf16: 3 rows in flight, as in the apply pass: 3.206 -> 3.189 ms/step on B200; TSS_BN_RED_U=4 for the A/B)
        static const int u_env = [] { const char* e = getenv("TSS_BN_RED_U"); return e ? atoi(e) : 3; }();
        const bool u3 = u_env != 4;
        int64_t want = ceil_div64(M, (int64_t)rl.PL * (u3 ? U3 : U) * 2);    // >= 2 batches of U rows per row lane
        const int64_t cap = (int64_t)tss_num_sms() * 3;           // the resident CTAs (<= 85 registers)
        const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
        auto kern = mask == kMaskNone ? (u3 ? bn_bwd_reduce_kernel<T, U3, kMaskNone> : bn_bwd_reduce_kernel<T, U, kMaskNone>)
                  : mask == kMaskZ    ? (u3 ? bn_bwd_reduce_kernel<T, U3, kMaskZ> : bn_bwd_reduce_kernel<T, U, kMaskZ>)
                                      : (u3 ? bn_bwd_reduce_kernel<T, U3, kMaskY> : bn_bwd_reduce_kernel<T, U, kMaskY>);
        tss_launch(kern, grid, rl.threads, 0, (cudaStream_t)stream,
            (const T*)dz, (const T*)z, (const T*)y, mean, rstd, gamma, beta, sums, M, C, lddz, ldz, ldy, rl.PL);
        TSS_LAUNCH_CHECK("bn_bwd_reduce");
        return TSS_OK;
    });
}

extern "C" int tss_bn_bwd_onepass(const void* dz, const void* z, const void* y, const float* mean, const float* rstd,
                                  const float* gamma, const float* beta, float* sums, void* dy, void* dres, float* dgamma,
                                  float* dbeta, int64_t M, int C, int64_t lddz, int64_t ldz, int64_t ldy, int64_t lddy,
                                  int64_t lddres, int flags, int* sync, int dtype, void* stream) {
    if (int e = check_rows("bn_bwd_onepass", M, C)) return e;
    TSS_REQUIRE(sync != nullptr && sums != nullptr, "bn_bwd_onepass: sums and sync are required");
    const int relu = flags & TSS_EPI_RELU;
    const RowsLaunch rl = rows_launch(M, C);
    TSS_REQUIRE(rl.threads > 0, "bn_bwd_onepass: C=%d too large", C);
    const int mask = !relu ? kMaskNone : (z != nullptr ? kMaskZ : kMaskY);
    TSS_DISPATCH_DTYPE(dtype, "bn_bwd_onepass", {
        constexpr int U = sizeof(T) == 2 ? 4 : 2;
        constexpr int UZ = sizeof(T) == 2 ? 2 : 1;               // three operands per row: fewer rows in flight
        const int u = mask == kMaskZ ? UZ : U;
        int64_t want = ceil_div64(M, (int64_t)rl.PL * u);
#ifdef TSS_HOST_EMU
        const int64_t cap = 1;                                    // blocks run one after the other on the emulation
#else
        // ALL CTAs resident at once: two per SM, minus slack for SMs that a kernel of another stream (NCCL) fills up
        const int64_t cap = (int64_t)tss_num_sms() * 2 - 40;
#endif
        const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
        auto kern = mask == kMaskNone ? bn_bwd_onepass_kernel<T, U, kMaskNone>
                  : mask == kMaskZ    ? bn_bwd_onepass_kernel<T, UZ, kMaskZ> : bn_bwd_onepass_kernel<T, U, kMaskY>;
        tss_launch(kern, grid, rl.threads, 0, (cudaStream_t)stream,
            (const T*)dz, (const T*)z, (const T*)y, mean, rstd, gamma, beta, sums, (T*)dy, (T*)dres, dgamma, dbeta, M, C,
            lddz, ldz, ldy, lddy, lddres, (float)(1.0 / (double)M), rl.PL, sync);
        TSS_LAUNCH_CHECK("bn_bwd_onepass");
        return TSS_OK;
    });
}

extern "C" int tss_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, const float* sums,
                                void* dy, void* dres, float* dgamma, float* dbeta, int64_t M, int64_t count, int C,
                                int64_t lddz, int64_t ldz, int64_t ldy, int64_t lddy, int64_t lddres, int flags,
                                int dtype, void* stream) {
    if (int e = check_rows("bn_bwd_apply", M, C)) return e;
    if (count <= 0) count = M;
    const int relu = flags & TSS_EPI_RELU;
    const RowsLaunch rl = rows_launch(M, C);
    TSS_REQUIRE(rl.threads > 0, "bn_bwd_apply: C=%d too large", C);
    TSS_DISPATCH_DTYPE(dtype, "bn_bwd_apply", {
        constexpr int U = sizeof(T) == 2 ? 4 : 2;
        constexpr int UZ = sizeof(T) == 2 ? 2 : 1;               // three operands per row (+ a second output): fewer rows in flight
        const int mask = !relu ? kMaskNone : (z != nullptr ? kMaskZ : kMaskY);
        // bf16: 3 rows in flight per operand fit the 80 registers of three resident CTAs; with 4 the compiler spills into the
        // loop (measured on B200 for the whole step: 3.286 ms with 4, 3.193 ms with 3; TSS_BN_BWD_U=4 for the A/B)
        constexpr int U3 = sizeof(T) == 2 ? 3 : 2;
        static const int u_env = [] { const char* e = getenv("TSS_BN_BWD_U"); return e ? atoi(e) : 3; }();
        const bool u3 = u_env != 4;
        auto kern = mask == kMaskNone ? (u3 ? bn_bwd_apply_kernel<T, U3, kMaskNone> : bn_bwd_apply_kernel<T, U, kMaskNone>)
                  : mask == kMaskZ    ? bn_bwd_apply_kernel<T, UZ, kMaskZ>
                                      : (u3 ? bn_bwd_apply_kernel<T, U3, kMaskY> : bn_bwd_apply_kernel<T, U, kMaskY>);
        static const int cap_env = [] { const char* e = getenv("TSS_BN_BWD_CAP"); return e ? atoi(e) : 8; }();   // CTAs per SM in the grid (A/B; 3 are resident)
        tss_launch(kern, rl.grid(mask == kMaskZ ? UZ : (u3 ? U3 : U), cap_env), rl.threads, 0, (cudaStream_t)stream,
            (const T*)dz, (const T*)z, (const T*)y, mean, rstd, gamma, beta, sums, (T*)dy, (T*)dres, dgamma, dbeta,
            M, C, lddz, ldz, ldy, lddy, lddres, (float)(1.0 / (double)count), rl.PL);
        TSS_LAUNCH_CHECK("bn_bwd_apply");
        return TSS_OK;
    });
}

extern "C" int tss_relu_bwd(const void* dz, const void* z, void* g, int64_t M, int C, int64_t lddz,
                            int64_t ldz, int64_t ldg, int dtype, void* stream) {
    if (int e = check_rows("relu_bwd", M, C)) return e;
    TSS_DISPATCH_DTYPE(dtype, "relu_bwd", {
        tss_launch(relu_bwd_kernel<T>, stream_grid(M * (C / 8)), kThreads, 0, (cudaStream_t)stream, 
            (const T*)dz, (const T*)z, (T*)g, M, C, lddz, ldz, ldg);
        TSS_LAUNCH_CHECK("relu_bwd");
        return TSS_OK;
    });
}

extern "C" int tss_add(const void* a, const void* b, void* out, int64_t M, int C, int64_t lda, int64_t ldb,
                       int64_t ldo, int dtype, void* stream) {
    if (int e = check_rows("add", M, C)) return e;
    TSS_DISPATCH_DTYPE(dtype, "add", {
        tss_launch(add_kernel<T>, stream_grid(M * (C / 8)), kThreads, 0, (cudaStream_t)stream, 
            (const T*)a, (const T*)b, (T*)out, M, C, lda, ldb, ldo);
        TSS_LAUNCH_CHECK("add");
        return TSS_OK;
    });
}

extern "C" int tss_copy_rows(const void* src, void* dst, int64_t M, int C, int64_t lds, int64_t ldd, int dtype,
                             void* stream) {
    if (int e = check_rows("copy_rows", M, C)) return e;
    TSS_DISPATCH_DTYPE(dtype, "copy_rows", {
        tss_launch(add_kernel<T>, stream_grid(M * (C / 8)), kThreads, 0, (cudaStream_t)stream, 
            (const T*)src, (const T*)nullptr, (T*)dst, M, C, lds, 0, ldd);
        TSS_LAUNCH_CHECK("copy_rows");
        return TSS_OK;
    });
}

extern "C" int tss_cast_from_f32(const float* src, void* dst, int64_t n, int dtype, void* stream) {
    TSS_REQUIRE(n > 0 && n % 8 == 0, "cast_from_f32: n=%lld must be a positive multiple of 8", (long long)n);
    TSS_DISPATCH_DTYPE(dtype, "cast_from_f32", {
        tss_launch(cast_from_f32_kernel<T>, stream_grid(n / 8), kThreads, 0, (cudaStream_t)stream, src, (T*)dst, n / 8);
        TSS_LAUNCH_CHECK("cast_from_f32");
        return TSS_OK;
    });
}

extern "C" int tss_scale_inplace(void* x, const float* s, int64_t n, int dtype, void* stream) {
    TSS_REQUIRE(n > 0 && n % 8 == 0, "scale_inplace: n=%lld must be a positive multiple of 8", (long long)n);
    TSS_DISPATCH_DTYPE(dtype, "scale_inplace", {
        tss_launch(scale_inplace_kernel<T>, stream_grid(n / 8), kThreads, 0, (cudaStream_t)stream, (T*)x, s, n / 8);
        TSS_LAUNCH_CHECK("scale_inplace");
        return TSS_OK;
    });
}
