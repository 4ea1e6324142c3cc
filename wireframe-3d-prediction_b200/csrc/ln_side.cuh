// Device-side bodies of the bf16 LayerNorm+ReLU passes between the tensor-core layers of the encoder
// (models/PointNetEncoder.py:38-39 and their backward), written so that the SAME code runs
//   * as a stand-alone kernel (ln_bf16.cu: 256 threads per CTA, the whole GPU), and
//   * as a "side job" of 128 spare threads inside the persistent tcgen05 GEMM kernel (gemm_tc.cu): the wide GEMMs are
//     bound by the tensor pipe and leave HBM almost idle, the LayerNorm passes are bound by HBM and leave the tensor pipe
//     idle -- run inside the GEMM of ANOTHER row chunk they cost (almost) no time of their own.
// The forward body is written over 256 VIRTUAL threads (NT real threads execute 256 / NT of them each), so its results --
// including the deterministic per-row-block column sums -- do not depend on NT: stand-alone and side runs are bit-identical.
#pragma once
#include "wf_common.cuh"

namespace wf {
namespace lnb {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// two bf16 in one 32-bit word -> (float(lo), float(hi))
__device__ __forceinline__ u64 bf2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u)); }
__device__ __forceinline__ uint32_t to_bf2(u64 v) {
    float lo, hi;
    up2(v, lo, hi);
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ u64 relu2(u64 v) { float lo, hi; up2(v, lo, hi); return pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }

__device__ __forceinline__ void load_pairs(const float* p, u64 (&f)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = pk2(a.x, a.y); f[1] = pk2(a.z, a.w); f[2] = pk2(b.x, b.y); f[3] = pk2(b.z, b.w);
}
__device__ __forceinline__ void load_pairs_smem(const float* p, u64 (&f)[4]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    f[0] = pk2(a.x, a.y); f[1] = pk2(a.z, a.w); f[2] = pk2(b.x, b.y); f[3] = pk2(b.z, b.w);
}

// h = relu(LN(z)) for one uint4 (8 channels) of a row with statistics (mu, rs)
__device__ __forceinline__ uint4 ln_relu8(const uint4& u, float mu, float rs, const u64 (&gm)[4], const u64 (&bt)[4]) {
    const u64 rs2 = pk2(rs, rs), nm2 = pk2(-mu * rs, -mu * rs);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = to_bf2(relu2(fma2(fma2(bf2(w[i]), rs2, nm2), gm[i], bt[i])));
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// (sum, sum of squares) partials of a row, added in part order -> (mean, rstd); ONE expression shared by wf_stats_finalize and the
// in-kernel finalisation of the GEMM's own-output LayerNorm, so that both give the same bits
__device__ __forceinline__ void stats_from_parts(const float2* __restrict__ st, size_t stride, int parts, float invC, float eps,
                                                 float& mu, float& rs) {
    float s1 = 0.f, s2 = 0.f;
    for (int p = 0; p < parts; ++p) { const float2 t = __ldcg(st + (size_t)p * stride); s1 += t.x; s2 += t.y; }     // fixed order
    mu = s1 * invC;
    const float var = fmaxf(s2 * invC - mu * mu, 0.f);
    rs = rsqrtf(var + eps);
}

// the same with a compile-time part count: all loads in flight at once (the in-kernel finalisation is latency-bound)
template <int PARTS>
__device__ __forceinline__ void stats_from_parts_n(const float2* __restrict__ st, size_t stride, float invC, float eps, float& mu, float& rs) {
    float2 t[PARTS];
#pragma unroll
    for (int p = 0; p < PARTS; ++p) t[p] = __ldcg(st + (size_t)p * stride);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int p = 0; p < PARTS; ++p) { s1 += t[p].x; s2 += t[p].y; }                                                    // fixed order
    mu = s1 * invC;
    const float var = fmaxf(s2 * invC - mu * mu, 0.f);
    rs = rsqrtf(var + eps);
}

// h = relu(LN(z)) on rows [row0, row0 + rows) of a [*, C8 * 8] bf16 tensor, statistics (mean, rstd) per local row in shared memory;
// NT threads, z read through L2 (ld.global.cg: the rows were written moments ago by other SMs of the same launch).  The slices
// are short (16-64 rows) and every load is an L2 round trip, so UN rows of every channel group are requested before any is used.
template <int C8, int NT>
__device__ __forceinline__ void ln_fwd_rows_l2(const uint4* __restrict__ z, uint4* __restrict__ h, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, const float* st, long long row0, int rows, int tid) {
    constexpr int G = C8 / NT, UN = 8;
    static_assert(C8 % NT == 0 && G >= 1 && G <= 2, "channel groups per thread");
    u64 gm[G][4], bt[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) { load_pairs(gamma + (tid + g * NT) * 8, gm[g]); load_pairs(beta + (tid + g * NT) * 8, bt[g]); }
    for (int r = 0; r < rows; r += UN) {
        uint4 u[G][UN];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < UN; ++j)
                u[g][j] = r + j < rows ? __ldcg(z + (row0 + r + j) * C8 + tid + g * NT) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < UN; ++j)
                if (r + j < rows) h[(row0 + r + j) * C8 + tid + g * NT] = ln_relu8(u[g][j], st[2 * (r + j)], st[2 * (r + j) + 1], gm[g], bt[g]);
    }
}

// barrier over the NT threads that execute a pass together: the whole CTA (stand-alone) or the side warps (named barrier)
template <int NT>
__device__ __forceinline__ void pass_sync(int bar_id) {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
}

// ------------------------------------------------------------------------------------------
// forward.  One call handles a block of CS_R consecutive rows.  256 virtual threads = (channel group c8, row phase vr);
// four rows' loads are issued before the first is used.  COLSUM (the LAST LayerNorm of the per-point MLP): also per-cloud
// column sums of the output h (all rows / rows with mask != 0).  The final Linear is affine, so the two mean pools of its
// output (models/PointNetEncoder.py:103-105, models/VertexPredictor.py:86) are that Linear applied to the mean of h:
// the (B,N,512) point-feature tensor never has to exist for them.  Deterministic: partial sums go to
// part[row block][segment][kind][C] (segment 1 = rows of the next cloud when the block straddles a cloud boundary) and are
// added in block order by seg_mean_kernel.
// ------------------------------------------------------------------------------------------
constexpr int CS_R = 128;

struct FwdArgs {
    const uint4* z; const float* mean; const float* rstd; const float* gamma; const float* beta; uint4* h;
    const uint8_t* mask; int M; int pool_n; int row_off; float* part;
};

// red: shared memory, (256 / C8) * 2 * C8 * 8 floats when COLSUM (unused otherwise)
template <int C8, bool COLSUM, int NT>
__device__ __forceinline__ void ln_fwd_block(const FwdArgs& a, int vblk, int tid, float* red, int bar_id) {
    constexpr int C = C8 * 8, VRS = 256 / C8, UN = 4, NV = 256 / NT;
    static_assert(NT == 128 || NT == 256, "thread count");
    const int blk_row0 = vblk * CS_R;
    const int rows = min(CS_R, a.M - blk_row0);
    const int g0 = a.row_off + blk_row0;                     // global row of this block's first row
    const int rb = COLSUM ? (g0 / a.pool_n + 1) * a.pool_n - g0 : rows;   // local rows >= rb belong to the next cloud
    const size_t gblk = (size_t)(g0 / CS_R);
    int c8[NV], vr[NV];
    u64 gm[NV][4], bt[NV][4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int vt = tid + v * NT;
        c8[v] = vt % C8; vr[v] = vt / C8;
        load_pairs(a.gamma + c8[v] * 8, gm[v]); load_pairs(a.beta + c8[v] * 8, bt[v]);
    }
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
        const int r_lo = seg == 0 ? 0 : rb, r_hi = seg == 0 ? min(rb, rows) : rows;
        if (r_lo >= r_hi) break;                             // uniform over the pass
        u64 su[NV][4], sm[NV][4];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int i = 0; i < 4; ++i) su[v][i] = sm[v][i] = 0ull;
        for (int base = r_lo; base < r_hi; base += UN * VRS) {
            uint4 u[NV][UN];
            float mu[NV][UN], rs[NV][UN];
            bool mk[NV][UN];
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int j = 0; j < UN; ++j) {
                    const int rr = base + vr[v] + j * VRS;
                    mk[v][j] = false; mu[v][j] = 0.f; rs[v][j] = 0.f; u[v][j] = make_uint4(0, 0, 0, 0);
                    if (rr < r_hi) {
                        const size_t row = (size_t)blk_row0 + rr;
                        u[v][j] = a.z[row * C8 + c8[v]]; mu[v][j] = a.mean[row]; rs[v][j] = a.rstd[row];
                        mk[v][j] = COLSUM && (a.mask == nullptr || a.mask[row] != 0);
                    }
                }
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int j = 0; j < UN; ++j) {
                    const int rr = base + vr[v] + j * VRS;
                    if (rr < r_hi) {
                        const uint4 o = ln_relu8(u[v][j], mu[v][j], rs[v][j], gm[v], bt[v]);
                        a.h[((size_t)blk_row0 + rr) * C8 + c8[v]] = o;
                        if (COLSUM) {                        // sum what the next GEMM will read (bf16-rounded)
                            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const u64 x = bf2(w[i]);
                                su[v][i] = add2(su[v][i], x);
                                if (mk[v][j]) sm[v][i] = add2(sm[v][i], x);
                            }
                        }
                    }
                }
        }
        if (COLSUM) {
            // red[vr][kind][c]
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float lo, hi;
                    float* r0 = red + ((size_t)vr[v] * 2 + 0) * C + c8[v] * 8 + 2 * i;
                    float* r1 = red + ((size_t)vr[v] * 2 + 1) * C + c8[v] * 8 + 2 * i;
                    up2(su[v][i], lo, hi); r0[0] = lo; r0[1] = hi;
                    up2(sm[v][i], lo, hi); r1[0] = lo; r1[1] = hi;
                }
            pass_sync<NT>(bar_id);
            float* dst = a.part + ((gblk * 2 + seg) * 2) * C;
            for (int i = tid; i < 2 * C; i += NT) {
                const int kind = i / C, c = i - kind * C;
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < VRS; ++q) t += red[((size_t)q * 2 + kind) * C + c];
                dst[(size_t)kind * C + c] = t;
            }
            pass_sync<NT>(bar_id);
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward, side-job form: NT = 128 threads, a thread owns G = C8 / NT groups of 8 consecutive channels; rows in groups of
// RG through an NSTG-stage cp.async ring in the side job's slice of the GEMM kernel's shared memory.  GB_SMEM: gain/shift
// live in shared memory instead of registers (the 2048-wide layer: two channel groups per thread would not fit otherwise).
//   g = dh * [y > 0];  gh = g * gamma;  c1 = mean(gh), c2 = mean(gh * xhat)
//   dz = rstd * (gh - c1 - xhat * c2);  dgamma += g * xhat;  dbeta += g;  dbias += dz
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

struct BwdArgs {
    const uint4* dh; const uint4* z; const float* mean; const float* rstd; const float* gamma; const float* beta;
    uint4* dz; float* dgamma; float* dbeta; float* dcolsum; long long M;
};

template <int C8, int NT, int RG, int NSTG, bool GB_SMEM>
constexpr int ln_bwd_side_smem() { return (GB_SMEM ? 2 * C8 * 8 * 4 : 0) + NSTG * 2 * RG * C8 * 16 + 2 * RG * NT * 4 + 2 * RG * 4 + 32; }

template <int C8, int NT, int RG, int NSTG, bool GB_SMEM>
__device__ __forceinline__ void ln_bwd_side(const BwdArgs& a, int worker, int n_workers, int tid, uint8_t* smem, int bar_id) {
    constexpr int C = C8 * 8, G = C8 / NT, NVAL = 2 * RG, NW = NT / 32;
    static_assert(C8 % NT == 0 && G >= 1 && G <= 2, "channel groups per thread");
    static_assert(NVAL % NW == 0 || NVAL < NW, "row sums per warp");
    float* gb = reinterpret_cast<float*>(smem);                                   // [2][C] when GB_SMEM
    uint4* ring = reinterpret_cast<uint4*>(smem + (GB_SMEM ? 2 * C * 4 : 0));      // [NSTG][2 (dh, z)][RG][C8]
    float* partial = reinterpret_cast<float*>(ring + NSTG * 2 * RG * C8);          // [NVAL][NT]
    float* rowsum = partial + NVAL * NT;                                           // [NVAL]
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
    u64 gm[GB_SMEM ? 1 : G][4], bt[GB_SMEM ? 1 : G][4];
    if (GB_SMEM) {
        for (int i = tid; i < C / 4; i += NT) {
            reinterpret_cast<float4*>(gb)[i] = __ldg(reinterpret_cast<const float4*>(a.gamma) + i);
            reinterpret_cast<float4*>(gb + C)[i] = __ldg(reinterpret_cast<const float4*>(a.beta) + i);
        }
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) { load_pairs(a.gamma + (tid + g * NT) * 8, gm[g]); load_pairs(a.beta + (tid + g * NT) * 8, bt[g]); }
    }
    u64 acc_g[G][4], acc_gx[G][4], acc_dz[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_g[g][i] = acc_gx[g][i] = acc_dz[g][i] = 0ull;
    const long long groups = (a.M + RG - 1) / RG;

    auto issue = [&](long long grp, int stage) {
        if (grp < groups) {
            const long long r0 = grp * RG;
#pragma unroll
            for (int r = 0; r < RG; ++r) {
                const bool ok = r0 + r < a.M;
                const long long row = ok ? r0 + r : 0;                      // size 0 -> zero fill, address stays valid
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int c8 = tid + g * NT;
                    const uint32_t d = ring_s + (((stage * 2 + 0) * RG + r) * C8 + c8) * 16;
                    cp_async16(d, a.dh + row * C8 + c8, ok ? 16u : 0u);
                    cp_async16(d + RG * C8 * 16, a.z + row * C8 + c8, ok ? 16u : 0u);
                }
            }
        }
        cp_async_commit();
    };
    auto load_stats = [&](long long g, float (&m)[RG], float (&s)[RG]) {
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const long long row = g * RG + r;
            const bool ok = g < groups && row < a.M;
            m[r] = ok ? __ldg(a.mean + row) : 0.f; s[r] = ok ? __ldg(a.rstd + row) : 0.f;
        }
    };

    long long grp = worker;
#pragma unroll
    for (int s = 0; s < NSTG - 1; ++s) issue(grp + (long long)s * n_workers, s);
    float mu_next[RG], rs_next[RG];
    load_stats(grp, mu_next, rs_next);
    if (GB_SMEM) pass_sync<NT>(bar_id);
    int stage = 0;
    for (; grp < groups; grp += n_workers) {
        issue(grp + (long long)(NSTG - 1) * n_workers, stage == 0 ? NSTG - 1 : stage - 1);
        float mu[RG], rs[RG];
#pragma unroll
        for (int r = 0; r < RG; ++r) { mu[r] = mu_next[r]; rs[r] = rs_next[r]; }
        load_stats(grp + n_workers, mu_next, rs_next);
        cp_async_wait<NSTG - 1>();                                           // this thread's copies of `grp` have landed
        const long long r0 = grp * RG;
        uint4* sd = ring + ((stage * 2 + 0) * RG) * C8;                      // thread reads/writes only its own 16-byte slots
        const uint4* sz = ring + ((stage * 2 + 1) * RG) * C8;
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const u64 rs2 = pk2(rs[r], rs[r]), nm2 = pk2(-mu[r] * rs[r], -mu[r] * rs[r]);
            u64 sa = 0ull, sb = 0ull;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c8 = tid + g * NT;
                const uint4 ud = sd[r * C8 + c8], uz = sz[r * C8 + c8];
                const uint32_t wd[4] = {ud.x, ud.y, ud.z, ud.w}, wz[4] = {uz.x, uz.y, uz.z, uz.w};
                u64 gmv[4], btv[4];
                if (GB_SMEM) { load_pairs_smem(gb + c8 * 8, gmv); load_pairs_smem(gb + C + c8 * 8, btv); }
                uint32_t gw[4];                                              // g = dh * [y > 0], still exact in bf16
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const u64 gmi = GB_SMEM ? gmv[i] : gm[GB_SMEM ? 0 : g][i], bti = GB_SMEM ? btv[i] : bt[GB_SMEM ? 0 : g][i];
                    const u64 xh = fma2(bf2(wz[i]), rs2, nm2);
                    float y0, y1;
                    up2(fma2(xh, gmi, bti), y0, y1);
                    gw[i] = wd[i] & ((y0 > 0.f ? 0x0000FFFFu : 0u) | (y1 > 0.f ? 0xFFFF0000u : 0u));
                    const u64 gh = mul2(bf2(gw[i]), gmi);
                    sa = add2(sa, gh); sb = fma2(gh, xh, sb);
                }
                sd[r * C8 + c8] = make_uint4(gw[0], gw[1], gw[2], gw[3]);    // the second pass reads g, not dh
            }
            float lo, hi;
            up2(sa, lo, hi); partial[(2 * r) * NT + tid] = lo + hi;
            up2(sb, lo, hi); partial[(2 * r + 1) * NT + tid] = lo + hi;
        }
        pass_sync<NT>(bar_id);
        for (int k = warp; k < NVAL; k += NW) {
            const float4 v = reinterpret_cast<const float4*>(partial + k * NT)[lane];   // NT = 128: four partials per lane
            float t = (v.x + v.y) + (v.z + v.w);
            t = warp_sum(t);
            if (lane == 0) rowsum[k] = t;
        }
        pass_sync<NT>(bar_id);
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            if (r0 + r >= a.M) break;
            const float c1 = rowsum[2 * r] * (1.0f / C), c2 = rowsum[2 * r + 1] * (1.0f / C);
            const u64 rs2 = pk2(rs[r], rs[r]), nm2 = pk2(-mu[r] * rs[r], -mu[r] * rs[r]);
            const u64 k1 = pk2(-c1 * rs[r], -c1 * rs[r]), k2 = pk2(-c2 * rs[r], -c2 * rs[r]);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c8 = tid + g * NT;
                const uint4 ug = sd[r * C8 + c8], uz = sz[r * C8 + c8];
                const uint32_t wg[4] = {ug.x, ug.y, ug.z, ug.w}, wz[4] = {uz.x, uz.y, uz.z, uz.w};
                u64 gmv[4];
                if (GB_SMEM) load_pairs_smem(gb + c8 * 8, gmv);
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const u64 gmi = GB_SMEM ? gmv[i] : gm[GB_SMEM ? 0 : g][i];
                    const u64 xh = fma2(bf2(wz[i]), rs2, nm2);
                    const u64 gg = bf2(wg[i]);
                    // dz = rstd * (g*gamma - c1 - xhat*c2)
                    const u64 dzv = fma2(xh, k2, fma2(mul2(gg, gmi), rs2, k1));
                    acc_g[g][i] = add2(acc_g[g][i], gg); acc_gx[g][i] = fma2(gg, xh, acc_gx[g][i]); acc_dz[g][i] = add2(acc_dz[g][i], dzv);
                    o[i] = to_bf2(dzv);
                }
                a.dz[(r0 + r) * C8 + c8] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        stage = stage + 1 == NSTG ? 0 : stage + 1;
    }
    cp_async_wait<0>();
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = (tid + g * NT) * 8 + 2 * i;
            float lo, hi;
            up2(acc_gx[g][i], lo, hi); atomicAdd(a.dgamma + c, lo); atomicAdd(a.dgamma + c + 1, hi);
            up2(acc_g[g][i], lo, hi); atomicAdd(a.dbeta + c, lo); atomicAdd(a.dbeta + c + 1, hi);
            up2(acc_dz[g][i], lo, hi); atomicAdd(a.dcolsum + c, lo); atomicAdd(a.dcolsum + c + 1, hi);
        }
    pass_sync<NT>(bar_id);                                                   // the shared-memory slice is free for the next segment
}

}  // namespace lnb
}  // namespace wf
