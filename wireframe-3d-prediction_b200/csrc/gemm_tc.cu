// wf_gemm_bf16: persistent, warp-specialised tcgen05 GEMM for the wide per-point encoder layers
// (models/PointNetEncoder.py:37-45, i=1..3 and the final projection) and their backward.
//
//   D[M,N] (+)= A * B^T (+bias)      bf16 operands, fp32 accumulation in TMEM
//
//   CTA tile 128 x 256, K step 64 (= one 128-byte swizzle row of bf16), 4-stage TMA ring
//   (48 KB per stage), two 256-column TMEM accumulator stages so the epilogue of tile i
//   overlaps the MMAs of tile i+1.  Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one
//   elected lane), warp 2 = TMEM allocator, warps 4..11 = epilogue (TMEM lane quarter = warp%4, two warps per quarter
//   splitting the tile's columns: with one warpgroup the K=512 layer was bound by the epilogue, not by the MMAs).
//
//   Operand layouts (both through the same 64-element-wide TMA boxes, SWIZZLE_128B):
//     K-major  : operand stored [rows, K], K contiguous  -> descriptor LBO unused, SBO = 1024 B
//     MN-major : operand stored [K, rows], rows contiguous (dW = dZ^T * H, reduction over
//                points) -> one (128 B x BK rows) box per 128 B of rows, LBO = box bytes, SBO = 1024 B
//   Element types: bf16 (kind::f16) for the encoder, tf32 on fp32 storage (kind::tf32, wf_gemm_tf32) for the
//   heads' fp32 row-MLPs -- same tile geometry in bytes, K per block = 64 resp. 32 elements.
//
//   Epilogue (thread = accumulator row): + bias, optional per-row (sum, sumsq) for the following
//   LayerNorm, then bf16 store, fp32 store or fp32 atomic accumulate (split-K).
#include "wf_common.cuh"
#include "sm100_ptx.cuh"
#include "ln_side.cuh"

#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace wf {
namespace tc {

// packed fp32 pairs (FADD2 / FFMA2): the epilogue's bias add and row statistics
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// One k-block is 128 bytes of K per row for either element type: 64 bf16 or 32 tf32 (fp32 storage).  All shared-memory
// byte sizes are therefore identical for both; only element counts differ.
constexpr int BM = 128, BN = 256, STAGES = 4;
constexpr int A_BYTES = BM * 128;             // 16 KB
constexpr int B_BYTES = BN * 128;             // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BAR_BYTES = 512;                                 // barriers + TMEM slot; keeps the staging boxes 512-byte aligned
constexpr int EPI_WARPS = 8;                                 // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int STG_TILE = 32 * 32;                            // floats: one 32x32 fp32 staging tile per epilogue warp, XOR-swizzled
constexpr int STG_BYTES = EPI_WARPS * STG_TILE * 4;
constexpr int BIAS_BYTES = EPI_WARPS * 32 * 4;                // per epilogue warp: the 32 bias values of the chunk in flight
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + STG_BYTES + BIAS_BYTES + 1024;
constexpr int TMEM_COLS = 512;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;
constexpr int MAXST = 6;                                      // most pipeline stages of any kernel form (barrier slots)
constexpr int SIDE_THREADS = 128;                             // SIDE kernels: warps 12..15 run the side-job segments
constexpr int SIDE_BAR = 1;                                   // named barrier of the side warps (0 = __syncthreads)
static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB per CTA)");

// staging tile addressing: element (row r, column c) of a 32x32 fp32 tile; 16-byte groups are XOR-swizzled with the row so
// that row-wise float4 writes (lane = row), row-segment float4 reads and column reads (lane = column) are all conflict-free
__device__ __forceinline__ int stg_v4(int r, int c4) { return (r * 8 + (c4 ^ (r & 7))) * 4; }

struct Params {
    int M, N, K;
    int tiles_m, tiles_n, split_k, kb_per_split, nkb;
    int streamk;                 // accumulate mode: balance (tile, k-block) units evenly over the workers
    const float* bias;
    void* D;
    int ldd, out_dtype, accumulate;
    int tma_store;               // bf16 output written by TMA stores through map_d (plain tiles: no accumulate / split / pool)
    long long split_stride;      // > 0: split s stores (not adds) its partial tile at D + s * split_stride (deterministic split-K)
    float* rowstats;
    // pool mode (final encoder Linear, models/PointNetEncoder.py:94,103-111 + models/VertexPredictor.py:87): instead of
    // storing D, every column's maximum over the rows of each cloud (pool_n consecutive rows) goes to a packed
    // 64-bit atomicMax -- (order-preserving bits of the value << 32) | ~(row within the cloud) -- so ties resolve to the
    // first row, exactly like torch.max(dim=1).  pool_max_u: all rows; pool_max_m: rows with mask != 0.
    int pool_n, pool_row0, pool_idx0;   // points per cloud (0 = off); global row index of this launch's row 0; offset added to
                                        // the stored point index (point-sharded clouds: this rank's first point)
    const uint8_t* pool_mask;
    unsigned long long* pool_max_u;
    unsigned long long* pool_max_m;
    // SIDE kernels: pipeline stages actually used (the ring's unused tail, (6 - n_stages) * 32 KB, is the side warps' shared
    // memory) and the side-job segments (include/wf_b200.h, wf_side_seg)
    int n_stages, n_side;
    // SIDE kernels, own-output LayerNorm (models/PointNetEncoder.py:38-39 fused behind the Linear of the SAME launch): the side
    // warps normalise each 256-row unit of D as soon as all its N tiles are stored -- the tile is still in L2, so the
    // LayerNorm pass reads nothing from DRAM.  own_done[m unit] counts the epilogue warps (2 CTAs x 8 per tile) whose stores
    // of that unit are complete; row statistics are finalised in-kernel from `rowstats` and also written out (backward).
    int own_kind;                // 0 = off, 1 = h = relu(LN(D)) (bf16)
    const float* own_gamma; const float* own_beta;
    void* own_h; float* own_mean; float* own_rstd; float own_eps;
    int* own_done;
    unsigned long long hint_a, hint_b, hint_d;   // L2 eviction hints of the TMA transfers (2-SM form)
    int n_workers;               // workers (CTAs, or CTA pairs) that take GEMM items; the rest of the grid only runs side jobs
    wf_side_seg side[WF_SIDE_MAX];
};

// float -> uint32 whose unsigned order is the float order (-inf < ... < -0 < +0 < ... < +inf)
__device__ __forceinline__ uint32_t ordered_bits(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void pool_push(unsigned long long* addr, float v, uint32_t n) {
    const unsigned long long pk = (static_cast<unsigned long long>(ordered_bits(v)) << 32) | (0xFFFFFFFFu - n);
    atomicMax(addr, pk);                              // result unused -> RED.MAX.64: fire-and-forget, no round trip to L2
}

// ESZ = operand element size: 2 -> bf16 (kind::f16), 4 -> tf32 on fp32 storage (kind::tf32).
// A_KM / B_KM: operand stored K-major ([rows, K]) or MN-major ([K, rows]).
// MC: CTAs run as clusters of 2 that work on the same N tile and adjacent M tiles in lock step; each CTA fetches half
//     of the shared B tile and TMA-multicasts it into both CTAs' shared memory (L2 -> SM traffic for B halves;
//     the kernel is L2-bandwidth-bound at 128x256 tiles otherwise).  MMAs stay cta_group::1.
// Work distribution, evaluated identically by the producer, MMA and epilogue roles.
//   classic : item w -> (tile, split); items are dealt round-robin to the workers (a worker = CTA, or cluster when MC)
//   stream-K: the (tile, k-block) space is cut into one contiguous, equally long range per worker; a range may end
//             inside a tile (its partial sum is added atomically), so the weight-gradient GEMMs -- few output tiles,
//             very long K -- keep every SM busy without wave quantisation.
struct Sched {
    int tiles_n, m_units, nkb, kb_per_split, n_items, w, w_step, streamk;
    long long u, u_end;
    __device__ __forceinline__ Sched(const Params& p, int m_units_, int worker, int n_workers)
        : tiles_n(p.tiles_n), m_units(m_units_), nkb(p.nkb), kb_per_split(p.kb_per_split),
          n_items(m_units_ * p.tiles_n * p.split_k), w(worker), w_step(n_workers), streamk(p.streamk) {
        const long long total = (long long)m_units_ * p.tiles_n * p.nkb;
        const long long per = (total + n_workers - 1) / n_workers;
        u = (long long)worker * per;
        u_end = u + per < total ? u + per : total;
        if (worker >= n_workers) { w = n_items; u = u_end = 0; }
    }
    __device__ __forceinline__ bool next(int& n_blk, int& mu, int& kb0, int& kb1) {
        if (streamk) {
            if (u >= u_end) return false;
            const long long tile = u / nkb;
            kb0 = (int)(u - tile * nkb);
            const long long take = (long long)(nkb - kb0) < u_end - u ? (long long)(nkb - kb0) : u_end - u;
            kb1 = kb0 + (int)take;
            u += take;
            n_blk = (int)(tile % tiles_n); mu = (int)(tile / tiles_n);
            return true;
        }
        if (w >= n_items) return false;
        n_blk = w % tiles_n; mu = (w / tiles_n) % m_units;
        const int split = w / (tiles_n * m_units);
        kb0 = split * kb_per_split; kb1 = min(nkb, kb0 + kb_per_split);
        w += w_step;
        return true;
    }
};

// ---- side jobs (warps 12..15 of a SIDE kernel): HBM-bound LayerNorm passes of another row chunk, see ln_side.cuh -------------
template <int C8, bool COLSUM>
__device__ __forceinline__ void side_ln_fwd(const wf_side_seg& sg, int tid, int worker, int n_workers, uint8_t* smem) {
    const lnb::FwdArgs a{static_cast<const uint4*>(sg.x0), sg.mean, sg.rstd, sg.gamma, sg.beta, static_cast<uint4*>(sg.out),
                         sg.mask, (int)sg.rows, COLSUM ? sg.pool_n : 1, COLSUM ? sg.row_off : 0, sg.part};
    const int nblk = (int)((sg.rows + lnb::CS_R - 1) / lnb::CS_R);
    for (int vb = worker; vb < nblk; vb += n_workers)
        lnb::ln_fwd_block<C8, COLSUM, SIDE_THREADS>(a, vb, tid, reinterpret_cast<float*>(smem), SIDE_BAR);
}
template <int C8, int RG, int NSTG, bool GB_SMEM>
__device__ __forceinline__ void side_ln_bwd(const wf_side_seg& sg, int tid, int worker, int n_workers, uint8_t* smem) {
    const lnb::BwdArgs a{static_cast<const uint4*>(sg.x0), static_cast<const uint4*>(sg.x1), sg.mean, sg.rstd, sg.gamma, sg.beta,
                         static_cast<uint4*>(sg.out), sg.acc0, sg.acc1, sg.acc2, (long long)sg.rows};
    lnb::ln_bwd_side<C8, SIDE_THREADS, RG, NSTG, GB_SMEM>(a, worker, n_workers, tid, smem, SIDE_BAR);
}
// shared memory a segment needs (host side: picks n_stages)
constexpr int SIDE_SMEM_FWD_COLSUM_1024 = 2 * 2 * 1024 * 4;
constexpr int SIDE_SMEM_BWD_1024 = lnb::ln_bwd_side_smem<128, SIDE_THREADS, 2, 3, false>();
constexpr int SIDE_SMEM_BWD_2048 = lnb::ln_bwd_side_smem<256, SIDE_THREADS, 2, 2, true>();
static_assert(SIDE_SMEM_FWD_COLSUM_1024 <= 32768 && SIDE_SMEM_BWD_1024 <= 32768 && SIDE_SMEM_BWD_2048 <= 65536, "side smem");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Own-output LayerNorm forward (Params::own_*): this CTA's side warps walk the CTA's own tile sequence; for every tile they
// wait until the whole 256-row unit it belongs to has been stored by all the CTAs that hold its N tiles, then normalise
// their slice of the unit's rows (256 rows / (2 * tiles_n) CTAs) over the full width N.
template <int C8>
__device__ __forceinline__ void own_ln_forward(const Params& p, int tid, uint8_t* smem, int m_units, int w_first, int crank) {
    Sched sched(p, m_units, w_first, p.n_workers);
    const int slices = 2 * p.tiles_n, R = (2 * BM) / slices;              // rows of a unit per CTA (16 .. 128)
    const int target = p.tiles_n * 2 * EPI_WARPS;
    float* st = reinterpret_cast<float*>(smem);                            // [R][2] mean, rstd  (2 * tiles_n = C8 / 16 row-statistic parts)
    const float invC = 1.0f / (float)p.N;
    int n_blk, mu, kb0, kb1;
    while (sched.next(n_blk, mu, kb0, kb1)) {
        const long long row0 = (long long)mu * (2 * BM) + (long long)(n_blk * 2 + crank) * R;
        const int rows = (int)(row0 + R <= p.M ? R : (p.M > row0 ? p.M - row0 : 0));
        if (tid == 0) {
            const int* flag = p.own_done + mu;
            if (ld_acquire_gpu(flag) < target) {
                const uint64_t t0 = ptx::globaltimer_ns();
                while (ld_acquire_gpu(flag) < target) {
                    __nanosleep(256);
                    if (ptx::globaltimer_ns() - t0 > WF_MBAR_TIMEOUT_NS) {
                        printf("wf_b200: own-output LayerNorm wait timed out (block %d unit %d: %d of %d)\n", (int)blockIdx.x, mu,
                               ld_acquire_gpu(flag), target);
                        __trap();
                    }
                }
            }
        }
        lnb::pass_sync<SIDE_THREADS>(SIDE_BAR);
        fence_proxy_async_global();                                        // D was written by the copy engines (async proxy)
        if (tid < rows) {
            float m_, r_;
            lnb::stats_from_parts_n<C8 / 16>(reinterpret_cast<const float2*>(p.rowstats) + (row0 + tid), (size_t)p.M, invC, p.own_eps, m_, r_);
            st[2 * tid] = m_; st[2 * tid + 1] = r_;
            p.own_mean[row0 + tid] = m_; p.own_rstd[row0 + tid] = r_;
        }
        lnb::pass_sync<SIDE_THREADS>(SIDE_BAR);
        if (rows > 0)
            lnb::ln_fwd_rows_l2<C8, SIDE_THREADS>(static_cast<const uint4*>(p.D), static_cast<uint4*>(p.own_h), p.own_gamma, p.own_beta,
                                                  st, row0, rows, tid);
        lnb::pass_sync<SIDE_THREADS>(SIDE_BAR);                            // st is rewritten for the next tile
    }
}

__device__ __forceinline__ void run_side_jobs(const Params& p, int tid, uint8_t* smem) {
    const int worker = blockIdx.x, n_workers = gridDim.x;
    for (int s = 0; s < p.n_side; ++s) {
        const wf_side_seg& sg = p.side[s];
        if (sg.kind == WF_SIDE_LN_FWD) {
            if (sg.C == 1024) side_ln_fwd<128, false>(sg, tid, worker, n_workers, smem);
            else side_ln_fwd<256, false>(sg, tid, worker, n_workers, smem);
        } else if (sg.kind == WF_SIDE_LN_FWD_COLSUM) {
            side_ln_fwd<128, true>(sg, tid, worker, n_workers, smem);
        } else if (sg.kind == WF_SIDE_LN_BWD) {
            if (sg.C == 1024) side_ln_bwd<128, 2, 3, false>(sg, tid, worker, n_workers, smem);
            else side_ln_bwd<256, 2, 2, true>(sg, tid, worker, n_workers, smem);
        }
    }
}

// MODE 0: one CTA per tile.  MODE 1: clusters of 2, cta_group::1 MMAs, B tile multicast (see MC above).
// MODE 2: clusters of 2 running ONE tcgen05.mma.cta_group::2 per k-step over a 256 x 256 tile: each CTA stages its own 128
//         rows of A and only HALF of the B tile (128 of the 256 N rows), so a k-block costs 32 KB instead of 48 KB of
//         shared-memory fill and read per CTA -- at 128x256 per CTA the 1-SM forms are bound by shared-memory bandwidth
//         (TMA writes + MMA operand reads ~188 B/clk of 128), the 2-SM form is not.  The leader CTA (rank 0) issues the
//         MMAs and owns the full/tmem-empty barriers; both CTAs' TMA bytes are credited to the leader's barrier; commits
//         are multicast so each CTA's producer and epilogue see their own barriers flip.  Six 32 KB stages.
// SIDE (MODE 2 only): 128 more threads (warps 12..15) execute the side-job segments of p.side while the other roles run the
//         GEMM; the ring uses p.n_stages (4 or 5) of its 6 slots and the rest is the side warps' shared memory; registers are
//         re-dealt with setmaxnreg (512 threads x 128 at launch: producer/MMA warpgroup down to 56, the two epilogue
//         warpgroups up to 160, the side warpgroup 136).
template <int ESZ, bool A_KM, bool B_KM, int MODE, bool SIDE>
__global__ void __launch_bounds__(NTHREADS + (SIDE ? SIDE_THREADS : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, const __grid_constant__ Params p) {
    constexpr bool MC = MODE == 1, TWO = MODE == 2, CL = MODE != 0;
    static_assert(!SIDE || TWO, "side jobs ride on the 2-SM kernel form");
    constexpr int NST_FULL = TWO ? 6 : STAGES;    // pipeline stages of the plain kernel forms
    constexpr int BB = TWO ? B_BYTES / 2 : B_BYTES;
    constexpr int SB = A_BYTES + BB;              // bytes per stage (NST_FULL * SB == STAGES * STAGE_BYTES)
    static_assert(NST_FULL * SB == STAGES * STAGE_BYTES && NST_FULL <= MAXST, "ring size");
    const int NST = SIDE ? p.n_stages : NST_FULL;
    constexpr int BK = 128 / ESZ;                 // elements of K per k-block
    constexpr int MNBOX = 128 / ESZ;              // MN-major: rows of the operand per TMA box (128 bytes)
    constexpr int BOX_BYTES = 128 * BK;           // MN-major box: BK k-rows x 128 bytes
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAXST + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAXST + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAXST + 2 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + STAGES * STAGE_BYTES + 8 * (2 * MAXST + 4));   // inside BAR_BYTES

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work items: clusters -> one item per CLUSTER covers two adjacent M tiles (this CTA takes 2*pair + rank)
    const int crank = CL ? (int)ptx::cluster_ctarank() : 0;
    const int m_units = CL ? (p.tiles_m + 1) / 2 : p.tiles_m;
    const int w_first = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = p.n_workers;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NST; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), MC ? 2 : 1); }
        // 2-SM: the leader's accumulator-free barrier collects the epilogue warps of both CTAs
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), TWO ? 2 * EPI_WARPS : EPI_WARPS); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if (TWO) { ptx::tmem_alloc2(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS); ptx::tmem_relinquish2(); }
        else     { ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (CL) ptx::cluster_sync();                 // peer barriers are initialised before any multicast / remote arrive
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // SIDE: the register file is re-dealt between the roles (whole warpgroups; 512 x 128 registers were allocated at launch).
    // Each role executes its own setmaxnreg at the top of its branch, so that the branch's code is compiled for that budget.
    if (SIDE && warp >= 12) {
        // ------------------------------------------------------------------ side jobs
        asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
        if (p.own_kind == 1) {
            if (p.N == 1024) own_ln_forward<128>(p, (int)threadIdx.x - NTHREADS, smem + NST * SB, m_units, w_first, crank);
            else own_ln_forward<256>(p, (int)threadIdx.x - NTHREADS, smem + NST * SB, m_units, w_first, crank);
        }
        run_side_jobs(p, (int)threadIdx.x - NTHREADS, smem + NST * SB);
    } else if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (SIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            Sched sched(p, m_units, w_first, w_step);
            int n_blk, mu, kb0, kb1;
            while (sched.next(n_blk, mu, kb0, kb1)) {
                const int m_blk = CL ? 2 * mu + crank : mu;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * SB, sb = sa + A_BYTES;
                    if (TWO) {
                        // own A rows and own half of the B rows; bytes of BOTH CTAs complete on the leader's barrier
                        const uint32_t lbar = full_bar(stage) & ptx::PEER_BIT_MASK;
                        if (crank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * SB);
                        if (A_KM) {
                            ptx::tma_load_2d_2sm_hint(sa, &map_a, lbar, kb * BK, m_blk * BM, p.hint_a);
                        } else {
#pragma unroll
                            for (int i = 0; i < BM / MNBOX; ++i)
                                ptx::tma_load_2d_2sm_hint(sa + i * BOX_BYTES, &map_a, lbar, m_blk * BM + i * MNBOX, kb * BK, p.hint_a);
                        }
                        if (B_KM) {
                            ptx::tma_load_2d_2sm_hint(sb, &map_b, lbar, kb * BK, n_blk * BN + crank * (BN / 2), p.hint_b);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / MNBOX / 2; ++i)
                                ptx::tma_load_2d_2sm_hint(sb + i * BOX_BYTES, &map_b, lbar, n_blk * BN + crank * (BN / 2) + i * MNBOX, kb * BK, p.hint_b);
                        }
                        if (++stage == NST) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    ptx::mbar_arrive_expect_tx(full_bar(stage), SB);
                    if (A_KM) {
                        ptx::tma_load_2d(sa, &map_a, full_bar(stage), kb * BK, m_blk * BM);
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / MNBOX; ++i)
                            ptx::tma_load_2d(sa + i * BOX_BYTES, &map_a, full_bar(stage), m_blk * BM + i * MNBOX, kb * BK);
                    }
                    if (!MC) {
                        if (B_KM) {
                            ptx::tma_load_2d(sb, &map_b, full_bar(stage), kb * BK, n_blk * BN);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / MNBOX; ++i)
                                ptx::tma_load_2d(sb + i * BOX_BYTES, &map_b, full_bar(stage), n_blk * BN + i * MNBOX, kb * BK);
                        }
                    } else {
                        // this CTA's half of the B tile, delivered to both CTAs (same smem offset, same barrier offset)
                        if (B_KM) {
                            ptx::tma_load_2d_mc(sb + crank * (B_BYTES / 2), &map_b, full_bar(stage), kb * BK,
                                                n_blk * BN + crank * (BN / 2), 3);
                        } else {
#pragma unroll
                            for (int i = 0; i < BN / MNBOX / 2; ++i) {
                                const int bi = crank * (BN / MNBOX / 2) + i;
                                ptx::tma_load_2d_mc(sb + bi * BOX_BYTES, &map_b, full_bar(stage), n_blk * BN + bi * MNBOX, kb * BK, 3);
                            }
                        }
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (SIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (lane == 0 && (!TWO || crank == 0)) {
            constexpr uint32_t idesc = ptx::idesc_f32acc(ESZ == 2 ? 1 : 2, TWO ? 2 * BM : BM, BN, A_KM ? 0 : 1, B_KM ? 0 : 1);
            // one tcgen05.mma consumes 32 bytes of K per row: K-major advances 32 B inside the swizzle row,
            // MN-major advances (32 / ESZ) k-rows of 128 B
            constexpr uint32_t LBO_A = A_KM ? 16u : (uint32_t)BOX_BYTES, LBO_B = B_KM ? 16u : (uint32_t)BOX_BYTES;
            constexpr uint32_t KSTEP_A = A_KM ? 32u : 128u * (32 / ESZ), KSTEP_B = B_KM ? 32u : 128u * (32 / ESZ);
            // MN-major tf32: swizzle atom is 4 k-rows of 128 B with 32-byte granularity (layout type 1), so the stride
            // between k-groups is 512 B; all other cases: 8 rows of 128 B (layout type 2), 1024 B
            constexpr uint32_t LT_A = (ESZ == 4 && !A_KM) ? 1u : 2u, LT_B = (ESZ == 4 && !B_KM) ? 1u : 2u;
            constexpr uint32_t SBO_A = (ESZ == 4 && !A_KM) ? 512u : 1024u, SBO_B = (ESZ == 4 && !B_KM) ? 512u : 1024u;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            Sched sched(p, m_units, w_first, w_step);
            int n_blk, mu, kb0, kb1;
            while (sched.next(n_blk, mu, kb0, kb1)) {
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + stage * SB, sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = ptx::smem_desc(sa + k * KSTEP_A, LBO_A, SBO_A, LT_A);
                        const uint64_t db = ptx::smem_desc(sb + k * KSTEP_B, LBO_B, SBO_B, LT_B);
                        const uint32_t accf = (kb > kb0 || k > 0) ? 1u : 0u;
                        if (TWO) { if (ESZ == 2) ptx::mma_f16_ss2(tmem_d, da, db, idesc, accf); else ptx::mma_tf32_ss2(tmem_d, da, db, idesc, accf); }
                        else     { if (ESZ == 2) ptx::mma_f16_ss(tmem_d, da, db, idesc, accf);  else ptx::mma_tf32_ss(tmem_d, da, db, idesc, accf); }
                    }
                    // frees the smem stage when the MMAs retire (clusters: in both CTAs)
                    if (TWO) ptx::mma_commit2_mc(empty_bar(stage), 3);
                    else if (MC) ptx::mma_commit_mc(empty_bar(stage), 3);
                    else ptx::mma_commit(empty_bar(stage));
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
                if (TWO) ptx::mma_commit2_mc(tfull_bar(acc), 3);  // accumulator halves ready for both CTAs' epilogues
                else ptx::mma_commit(tfull_bar(acc));             // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < 4) {
        if (SIDE) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");     // TMEM-allocator and spare warp: same warpgroup
    } else if (warp < 12) {
        // ------------------------------------------------------------------ epilogue
        if (SIDE) asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
        // TMEM -> registers (thread = row): + bias, row statistics.  Then through a per-warp 32x32 fp32 staging tile in
        // shared memory so that global stores are row-contiguous (a quarter warp writes one 128-byte fp32 row segment /
        // a 64-byte bf16 segment) instead of 32 rows x 16 bytes per instruction.
        const int q = warp & 3;                                  // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;                        // which half of the tile's column chunks this warp owns
        float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BAR_BYTES) + (warp - 4) * STG_TILE;
        float* bias_slot = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BAR_BYTES + STG_BYTES) + (warp - 4) * 32;
        int acc = 0; uint32_t acc_phase = 0;
        int own_prev = -1;                                       // own-output LayerNorm: unit whose completion is still to be signalled
        Sched sched(p, m_units, w_first, w_step);
        int n_blk, mu, kb0, kb1;
        // pool mode: the validity byte of this thread's row is fetched ONE TILE AHEAD (a second scheduler walks one item in front):
        // read at the top of its own tile, the load's latency was exposed once per tile -- 11 % of the pooling launch's stall
        // samples (profiles/r02_pool_source_top.txt)
        Sched ahead(p, m_units, w_first, w_step);
        auto mask_of = [&](int mu_) -> uint32_t {
            const int r_ = (CL ? 2 * mu_ + crank : mu_) * BM + q * 32 + lane;
            return (r_ < p.M && (p.pool_mask == nullptr || p.pool_mask[r_] != 0)) ? 1u : 0u;
        };
        uint32_t mk_next = 0;
        bool has_next = false;
        if (p.pool_n > 0) {
            int an, amu, ak0, ak1;
            if (ahead.next(an, amu, ak0, ak1)) mk_next = mask_of(amu);          // the first item's own byte
            has_next = true;
        }
        while (sched.next(n_blk, mu, kb0, kb1)) {
            const int m_blk = CL ? 2 * mu + crank : mu;
            const int row0 = m_blk * BM + q * 32;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            // pool mode: which of this warp's 32 rows exist / are valid points, and where the next cloud starts
            uint32_t pool_mbits = 0; int pool_rows = 0, pool_b0 = 0, pool_rb = 32;
            if (p.pool_n > 0) {
                const bool mk = mk_next != 0;
                if (has_next) {
                    int an, amu, ak0, ak1;
                    has_next = ahead.next(an, amu, ak0, ak1);
                    if (has_next) mk_next = mask_of(amu);
                }
                pool_mbits = __ballot_sync(0xffffffffu, mk);
                pool_rows = min(32, p.M - row0);
                pool_b0 = (p.pool_row0 + row0) / p.pool_n;
                pool_rb = (pool_b0 + 1) * p.pool_n - (p.pool_row0 + row0);   // rows >= pool_rb belong to the next cloud
            }
            if (SIDE && p.own_kind != 0 && own_prev >= 0) {
                // Tell the side warps that this warp's part of the PREVIOUS tile's unit is in global memory.  Done here, one
                // tile late and before the wait for the next accumulator: the bulk stores and the row statistics of that tile
                // were issued a whole tile ago, so neither the wait_group nor the fence has anything left to wait for.
                __syncwarp();                                    // orders the other lanes' row-statistics stores before lane 0's release
                if (lane == 0) {
                    ptx::bulk_wait<0>();
                    fence_proxy_async_global();
                    __threadfence();
                    atomicAdd(p.own_done + own_prev, 1);
                }
                own_prev = -1;
            }
            const bool add_bias = p.bias != nullptr && kb0 == 0;       // split-K: the first K range carries the bias
            // Bias: lane j fetches the value of column j of the NEXT chunk one chunk ahead (one register), so the load's
            // latency hides behind the accumulator wait / the previous chunk; it reaches all rows (= threads) through a
            // 128-byte per-warp slot of shared memory read back as broadcast float4s.  Columns beyond N get 0.
            const int c_first = half * (BN / 64), c_end = (half + 1) * (BN / 64);
            float bias_next = 0.f;
            if (add_bias) { const int col = n_blk * BN + c_first * 32 + lane; if (col < p.N) bias_next = __ldg(p.bias + col); }
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            // deterministic split-K: every K range has its own fp32 slice of the workspace, summed in order afterwards
            void* const Dt = p.split_stride > 0 ? static_cast<void*>(static_cast<float*>(p.D) + (long long)(kb0 / p.kb_per_split) * p.split_stride)
                                                : p.D;
            u64 s1a = 0ull, s1b = 0ull, s2a = 0ull, s2b = 0ull;   // row (sum, sum of squares): two packed partial chains each
            float s1 = 0.f, s2 = 0.f;                              // column-tail chunks
#pragma unroll 1
            for (int c = c_first; c < c_end; ++c) {
                const int col0 = n_blk * BN + c * 32;
                if (col0 >= p.N) break;                          // warp-uniform
                if (add_bias) {
                    __syncwarp();
                    bias_slot[lane] = bias_next;
                    bias_next = 0.f;
                    if (c + 1 < c_end) { const int col = col0 + 32 + lane; if (col < p.N) bias_next = __ldg(p.bias + col); }
                    __syncwarp();
                }
                uint32_t r[32];
                ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, r);
                const bool full = col0 + 32 <= p.N;
                float4 bb[8];
                if (add_bias) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) bb[j] = reinterpret_cast<const float4*>(bias_slot)[j];
                }
                ptx::tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (add_bias) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        up2(add2(pk2(v[4 * j], v[4 * j + 1]), pk2(bb[j].x, bb[j].y)), v[4 * j], v[4 * j + 1]);
                        up2(add2(pk2(v[4 * j + 2], v[4 * j + 3]), pk2(bb[j].z, bb[j].w)), v[4 * j + 2], v[4 * j + 3]);
                    }
                }
                if (p.rowstats != nullptr) {
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const u64 a = pk2(v[j], v[j + 1]), b = pk2(v[j + 2], v[j + 3]);
                            s1a = add2(s1a, a); s2a = fma2(a, a, s2a);
                            s1b = add2(s1b, b); s2b = fma2(b, b, s2b);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
                    }
                }
                if (p.pool_n > 0) {
                    // ---- stage, then lane = column: running max / first argmax over this warp's rows, per cloud
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(stg + stg_v4(lane, j >> 2)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    __syncwarp();
                    const int col = col0 + lane;
                    if (pool_rb >= 32 && pool_rows == 32) {
                        // ---- all 32 rows exist and belong to one cloud (the usual tile): fully unrolled walk, two independent
                        // chains (rows 0..15 / 16..31) per kind, merged with a strict comparison (ties keep the earlier row).
                        // The generic loop below costs ~18 instructions per row (loop, swizzle arithmetic, "first row" tests) in
                        // one dependent chain; at K = 1024 that made the epilogue longer than the tile's MMAs (tensor pipe 60 %).
                        const float* sp = stg + (lane & 3);
                        const int c4 = lane >> 2;
                        float mu0 = sp[(0 * 8 + (c4 ^ 0)) * 4], mu1 = sp[(16 * 8 + (c4 ^ 0)) * 4];
                        int au0 = 0, au1 = 16;
                        float mm0 = 0.f, mm1 = 0.f; int am0 = -1, am1 = -1;
                        if (pool_mbits == 0xFFFFFFFFu || pool_mbits == 0u) {
                            // every row is a valid point (all tiles of a cloud but its zero-padded tail): the masked maximum IS the
                            // unmasked one -- one chain pair instead of two (4 instead of 10 instructions per row); no row valid
                            // (inside the padded tail): there is no masked maximum
#pragma unroll
                            for (int r = 1; r < 16; ++r) {
                                const float t0 = sp[(r * 8 + (c4 ^ (r & 7))) * 4], t1 = sp[((r + 16) * 8 + (c4 ^ (r & 7))) * 4];
                                if (t0 > mu0) { mu0 = t0; au0 = r; }
                                if (t1 > mu1) { mu1 = t1; au1 = r + 16; }
                            }
                            if (mu1 > mu0) { mu0 = mu1; au0 = au1; }
                            if (pool_mbits != 0u) { mm0 = mu0; am0 = au0; }
                        } else {
                            if (pool_mbits & 1u) { mm0 = mu0; am0 = 0; }
                            if ((pool_mbits >> 16) & 1u) { mm1 = mu1; am1 = 16; }
#pragma unroll
                            for (int r = 1; r < 16; ++r) {
                                const float t0 = sp[(r * 8 + (c4 ^ (r & 7))) * 4], t1 = sp[((r + 16) * 8 + (c4 ^ (r & 7))) * 4];
                                if (t0 > mu0) { mu0 = t0; au0 = r; }
                                if (t1 > mu1) { mu1 = t1; au1 = r + 16; }
                                if (((pool_mbits >> r) & 1u) && (am0 < 0 || t0 > mm0)) { mm0 = t0; am0 = r; }
                                if (((pool_mbits >> (r + 16)) & 1u) && (am1 < 0 || t1 > mm1)) { mm1 = t1; am1 = r + 16; }
                            }
                            if (mu1 > mu0) { mu0 = mu1; au0 = au1; }
                            if (am1 >= 0 && (am0 < 0 || mm1 > mm0)) { mm0 = mm1; am0 = am1; }
                        }
                        if (col < p.N && p.pool_max_u != nullptr) {        // (nullptr: WF_B200_POOL_NOPUSH diagnosis, see the launcher)
                            const int base = p.pool_idx0 + p.pool_row0 + row0 - pool_b0 * p.pool_n;
                            const size_t o = (size_t)pool_b0 * p.N + col;
                            pool_push(p.pool_max_u + o, mu0, (uint32_t)(base + au0));
                            if (am0 >= 0) pool_push(p.pool_max_m + o, mm0, (uint32_t)(base + am0));
                        }
                    } else
#pragma unroll 1
                    for (int seg = 0; seg < 2; ++seg) {
                        const int r_lo = seg == 0 ? 0 : pool_rb;
                        const int r_hi = seg == 0 ? min(pool_rb, pool_rows) : pool_rows;
                        if (r_lo >= r_hi) continue;                          // warp-uniform
                        float mu = 0.f, mm = 0.f; int au = -1, am = -1;
#pragma unroll 4
                        for (int r = r_lo; r < r_hi; ++r) {
                            const float t = stg[stg_v4(r, lane >> 2) + (lane & 3)];
                            if (au < 0 || t > mu) { mu = t; au = r; }
                            if (((pool_mbits >> r) & 1u) && (am < 0 || t > mm)) { mm = t; am = r; }
                        }
                        if (col < p.N && p.pool_max_u != nullptr) {
                            const int b = pool_b0 + seg;
                            const int base = p.pool_idx0 + p.pool_row0 + row0 - b * p.pool_n;            // row index inside the cloud of local row 0
                            const size_t o = (size_t)b * p.N + col;
                            pool_push(p.pool_max_u + o, mu, (uint32_t)(base + au));
                            if (am >= 0) pool_push(p.pool_max_m + o, mm, (uint32_t)(base + am));
                        }
                    }
                    __syncwarp();
                } else if (p.tma_store) {
                    // ---- bf16 through TMA: the thread (= row) converts its 32 columns and writes them as one 64-byte row of a
                    // 32 x 32 box in the SWIZZLE_64B layout (16-byte unit ^= address bits 7-8: conflict-free for row-wise
                    // writes); one lane hands the box to the copy engine, which clips rows >= M and columns >= N.  Two boxes
                    // per warp alternate, so the warp never waits for a store it has just issued -- no read-back of the
                    // staged tile, no per-thread global stores.
                    const int buf = (c - c_first) & 1;
                    if (lane == 0) ptx::bulk_wait_read<1>();              // the box written two chunks ago has been read
                    __syncwarp();
                    const uint32_t box_off = (uint32_t)(STAGES * STAGE_BYTES + BAR_BYTES + (warp - 4) * (STG_TILE * 4) + buf * 2048);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint4 pk;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * k], v[8 * k + 1]), t1 = __floats2bfloat162_rn(v[8 * k + 2], v[8 * k + 3]);
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * k + 4], v[8 * k + 5]), t3 = __floats2bfloat162_rn(v[8 * k + 6], v[8 * k + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                        uint32_t off = box_off + lane * 64 + k * 16;
                        off ^= ((off >> 7) & 3u) << 4;
                        *reinterpret_cast<uint4*>(smem + off) = pk;
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d_hint(&map_d, smem_base + box_off, col0, row0, p.hint_d);
                        ptx::bulk_commit();
                    }
                } else if (full && !p.accumulate) {
                    // ---- stage, then row-contiguous stores
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(stg + stg_v4(lane, j >> 2)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    __syncwarp();
                    if (p.out_dtype == WF_BF16) {
                        const int rr = lane >> 2, cc = (lane & 3) * 8;         // 4 lanes per row, 8 columns each
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rl = rr + 8 * i;
                            const float4 a = *reinterpret_cast<const float4*>(stg + stg_v4(rl, cc >> 2));
                            const float4 b = *reinterpret_cast<const float4*>(stg + stg_v4(rl, (cc >> 2) + 1));
                            if (row0 + rl < p.M) {
                                uint4 pk;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x, a.y), t1 = __floats2bfloat162_rn(a.z, a.w);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(b.x, b.y), t3 = __floats2bfloat162_rn(b.z, b.w);
                                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(Dt) + (size_t)(row0 + rl) * p.ldd + col0 + cc) = pk;
                            }
                        }
                    } else {
                        const int rr = lane >> 3, cc = (lane & 7) * 4;         // 8 lanes per row, 4 columns each
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rl = rr + 4 * i;
                            const float4 a = *reinterpret_cast<const float4*>(stg + stg_v4(rl, cc >> 2));
                            if (row0 + rl < p.M)
                                *reinterpret_cast<float4*>(static_cast<float*>(Dt) + (size_t)(row0 + rl) * p.ldd + col0 + cc) = a;
                        }
                    }
                    __syncwarp();
                } else if (row_ok) {
                    // ---- column tail or split-K accumulation: direct per-row access
                    if (p.out_dtype == WF_BF16) {
                        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(Dt) + (size_t)row * p.ldd + col0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)                      // static indices: keeps v[] in registers
                            if (col0 + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
                    } else {
                        float* dst = static_cast<float*>(Dt) + (size_t)row * p.ldd + col0;
                        if (p.accumulate) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (full || col0 + j < p.N) atomicAdd(dst + j, v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.N) dst[j] = v[j];
                        }
                    }
                }
            }
            if (p.rowstats != nullptr && row_ok) {   // one slot per (N tile, row): summed in tile order by wf_stats_finalize
                float a0, a1, b0, b1;
                up2(add2(s1a, s1b), a0, a1); up2(add2(s2a, s2b), b0, b1);
                reinterpret_cast<float2*>(p.rowstats)[(size_t)(2 * n_blk + half) * p.M + row] = make_float2((a0 + a1) + s1, (b0 + b1) + s2);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (TWO) ptx::mbar_arrive_cluster(tempty_bar(acc) & ptx::PEER_BIT_MASK);     // the leader's barrier
                else ptx::mbar_arrive(tempty_bar(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            if (SIDE && p.own_kind != 0) own_prev = mu;
        }
        if (p.tma_store && lane == 0) ptx::bulk_wait<0>();       // this warp's stores are complete before the CTA may exit
        if (SIDE && p.own_kind != 0 && own_prev >= 0) {
            __syncwarp();
            if (lane == 0) { fence_proxy_async_global(); __threadfence(); atomicAdd(p.own_done + own_prev, 1); }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (CL) ptx::cluster_sync();                 // nobody exits while the peer may still write our smem / barriers
    if (warp == 2) { if (TWO) ptx::tmem_dealloc2(tmem_base, TMEM_COLS); else ptx::tmem_dealloc(tmem_base, TMEM_COLS); }
}

// out[m][n] = bias[n] + sum_s part[s][m][n], s in order: the second half of the deterministic split-K
__global__ void splitk_reduce_kernel(const float4* __restrict__ part, int split, long long mn4, int n4, const float4* __restrict__ bias,
                                     float* __restrict__ out, int ldo, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < mn4; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / n4; const int c4 = (int)(i - m * n4);
        float4 a = bias ? bias[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < split; ++s) { const float4 t = part[(long long)s * mn4 + i]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
        float4* dst = reinterpret_cast<float4*>(out + m * ldo + 4 * c4);
        if (accumulate) { const float4 o = *dst; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
        *dst = a;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// 2-D tensor map: inner dimension `inner` elements (contiguous), `outer` rows of stride ld elements.
// MN-major fp32 (tf32) operands need the 32-byte-atom flavour of the 128-byte swizzle (UMMA SWIZZLE_128B_BASE32B);
// every other case uses the plain 128-byte swizzle.
static int make_map(CUtensorMap* m, int esz, bool mn_major, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                    uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return WF_ECUDA; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * (uint64_t)esz};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                     const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (esz == 4 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld); return WF_ECUDA; }
    return WF_OK;
}

// bf16 output [outer rows, inner columns], boxes of 32 x 32 elements (64-byte rows) in the 64-byte swizzle
static int make_store_map(CUtensorMap* m, void* ptr, uint64_t inner, uint64_t outer, uint64_t ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return WF_ECUDA; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2ull};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (store map) failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld); return WF_ECUDA; }
    return WF_OK;
}

}  // namespace tc
}  // namespace wf

namespace wf { namespace tc {

template <int ESZ, bool A_KM, bool B_KM, int MC, bool SIDE>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const Params& p, int grid, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gemm_tc_kernel<ESZ, A_KM, B_KM, MC, SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    });
    WF_CUDA(attr_err);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS + (SIDE ? SIDE_THREADS : 0)); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MC != 0 ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    WF_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<ESZ, A_KM, B_KM, MC, SIDE>, ma, mb, md, p));
    return WF_OK;
}

template <int ESZ, bool A_KM, bool B_KM>
static int launch_mc(int mode, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const Params& p, int grid, cudaStream_t s) {
    if (mode == 2) return launch<ESZ, A_KM, B_KM, 2, false>(ma, mb, md, p, grid, s);
    if (mode == 1) return launch<ESZ, A_KM, B_KM, 1, false>(ma, mb, md, p, grid, s);
    return launch<ESZ, A_KM, B_KM, 0, false>(ma, mb, md, p, grid, s);
}

// WF_B200_GEMM_MODE: 2 (default) = 2-SM MMA, 1 = 1-SM MMA with multicast B, 0 = no clusters
static int cluster_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("WF_B200_GEMM_MODE"); v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }
    return v;
}

struct PoolArgs { int n, row0, idx0; const uint8_t* mask; unsigned long long* max_u; unsigned long long* max_m; };
struct OwnArgs { const float* gamma; const float* beta; void* h; float* mean; float* rstd; float eps; int* done; };

// esz 2: bf16 operands; esz 4: fp32 operands multiplied as tf32
static int gemm_tc(int esz, const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, int M, int N, int K,
                   const float* bias, void* D, int ldd, int out_dtype, int accumulate, int split_k, float* rowstats,
                   cudaStream_t stream, const PoolArgs* pool = nullptr, float* det_work = nullptr, long long det_work_floats = 0,
                   const wf_side_seg* segs = nullptr, int n_segs = 0, const OwnArgs* own = nullptr) {
    const int al = 16 / esz;                                     // elements per 16 bytes
    WF_CHECK_ARG(M > 0 && N > 0 && K > 0, "wf_gemm_tc: empty problem M=%d N=%d K=%d", M, N, K);
    WF_CHECK_ARG(lda % al == 0 && ldb % al == 0, "wf_gemm_tc: lda/ldb must be multiples of %d elements (16-byte TMA strides)", al);
    WF_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                 "wf_gemm_tc: operands must be 16-byte aligned");
    WF_CHECK_ARG(out_dtype == WF_BF16 || out_dtype == WF_F32, "wf_gemm_tc: bad out_dtype");
    WF_CHECK_ARG(!(accumulate && out_dtype != WF_F32), "wf_gemm_tc: accumulate needs an fp32 output");
    WF_CHECK_ARG(out_dtype == WF_F32 ? (ldd % 4 == 0) : (ldd % 8 == 0), "wf_gemm_tc: ldd alignment");
    WF_CHECK_ARG((reinterpret_cast<uintptr_t>(D) & 15) == 0, "wf_gemm_tc: D must be 16-byte aligned");
    WF_CHECK_ARG(pool != nullptr || D != nullptr, "wf_gemm_tc: D is null");
    if (split_k < 1) split_k = 1;
    WF_CHECK_ARG(split_k == 1 || accumulate || det_work != nullptr, "wf_gemm_tc: split_k > 1 needs accumulate or a workspace");
    const int BKe = 128 / esz, mnbox = 128 / esz;
    CUtensorMap ma, mb;
    int rc;
    if (a_kmajor) { if ((rc = make_map(&ma, esz, false, A, K, M, lda, BKe, BM)) != WF_OK) return rc; }
    else          { if ((rc = make_map(&ma, esz, true, A, M, K, lda, mnbox, BKe)) != WF_OK) return rc; }
    // multicast pairs pay off when there are at least two M tiles to pair up
    const int mode = (cdiv(M, BM) >= 2 && sm_count() >= 2) ? cluster_mode() : 0;
    const bool mc = mode != 0;
    if (n_segs > 0) {
        WF_CHECK_ARG(mode == 2 && esz == 2 && a_kmajor == b_kmajor, "wf_gemm_bf16_side: side jobs need the 2-SM bf16 kernel form "
                     "(M >= 256, WF_B200_GEMM_MODE=2, both operands K-major or both MN-major)");
        WF_CHECK_ARG(segs != nullptr && n_segs <= WF_SIDE_MAX, "wf_gemm_bf16_side: at most %d segments", WF_SIDE_MAX);
    }
    if (b_kmajor) { if ((rc = make_map(&mb, esz, false, B, K, N, ldb, BKe, mc ? BN / 2 : BN)) != WF_OK) return rc; }
    else          { if ((rc = make_map(&mb, esz, true, B, N, K, ldb, mnbox, BKe)) != WF_OK) return rc; }
    Params p;
    p.M = M; p.N = N; p.K = K;
    p.tiles_m = cdiv(M, BM); p.tiles_n = cdiv(N, BN);
    p.nkb = cdiv(K, BKe);
    const int m_units = mc ? (p.tiles_m + 1) / 2 : p.tiles_m;
    const int workers_max = mc ? sm_count() / 2 : sm_count();
    static const bool streamk_env = [] { const char* e = getenv("WF_B200_GEMM_STREAMK"); return e && e[0] == '1'; }();
    p.streamk = (streamk_env && accumulate && split_k > 1 && bias == nullptr && rowstats == nullptr) ? 1 : 0;
    // WF_B200_DW_ALIGN=1 (default off): split-K products with fewer output tiles than workers (the weight gradients: 8-32
    // tiles, K = all points) on groups * tiles workers (64 of 74 pairs for 32 tiles) with a split count that is a multiple of
    // the groups, so that a K range never straddles two waves.  Tried against the 1.4-1.8x DRAM reads of the weight gradients:
    // the reads did not change (5.6 -> 5.5 GB, profiles/r02_gemm_traffic_*) -- the concurrent readers of a K range drift apart
    // by more than L2 holds -- and neither did the time.
    int group_workers = 0;
    static const bool dw_align = [] { const char* e = getenv("WF_B200_DW_ALIGN"); return e && e[0] == '1'; }();
    if (dw_align && accumulate && split_k > 1 && !p.streamk && det_work == nullptr) {
        const long long tiles = (long long)m_units * p.tiles_n;
        const int groups = (int)(workers_max / (tiles > 0 ? tiles : 1));
        if (groups >= 1 && groups * tiles * 5 >= (long long)workers_max * 4) {       // at least 80 % of the workers stay busy
            group_workers = (int)(groups * tiles);
            split_k = ((split_k + groups - 1) / groups) * groups;
        }
    }
    if (split_k > 1 && !p.streamk && group_workers == 0) {
        // The caller's split_k is a hint ("reduction-heavy").  Items are dealt round-robin, split index slowest, so the
        // CTAs running together read the same K range of both operands (L2 reuse); pick the smallest split whose
        // item count fills whole waves of workers (>= 95 %), which is what removes the wave-quantisation loss.
        const long long tiles = (long long)m_units * p.tiles_n;
        int best = 1; double best_eff = 0.0;
        int smax = p.nkb / 4 < 64 ? (p.nkb / 4 < 1 ? 1 : p.nkb / 4) : 64;
        if (det_work != nullptr && smax > split_k) smax = split_k;       // the caller sized the workspace for its own split count
        for (int sp = 1; sp <= smax; ++sp) {
            const long long it = tiles * sp;
            const long long waves = (it + workers_max - 1) / workers_max;
            const double eff = (double)it / (double)(waves * workers_max);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
            if (eff >= 0.95) { best = sp; break; }
        }
        split_k = best;
    }
    if (split_k > p.nkb) split_k = p.nkb;
    p.kb_per_split = cdiv(p.nkb, split_k);
    p.split_k = cdiv(p.nkb, p.kb_per_split);
    p.bias = bias; p.D = D; p.ldd = ldd; p.out_dtype = out_dtype; p.accumulate = accumulate; p.rowstats = rowstats;
    p.split_stride = 0;
    const bool det = det_work != nullptr && p.split_k > 1;
    if (det) {
        // deterministic split-K: partial tiles go to [split][M][N] fp32 slices of the workspace, splitk_reduce_kernel adds them
        WF_CHECK_ARG(out_dtype == WF_F32 && N % 4 == 0 && ldd % 4 == 0 && rowstats == nullptr && pool == nullptr,
                     "wf_gemm_tc: deterministic split-K needs an fp32 output with N %% 4 == 0");
        WF_CHECK_ARG((long long)p.split_k * M * N <= det_work_floats,
                     "wf_gemm_tc: split-K workspace too small (%lld floats needed)", (long long)p.split_k * M * N);
        WF_CHECK_ARG((reinterpret_cast<uintptr_t>(det_work) & 15) == 0 && (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
                     "wf_gemm_tc: workspace / bias alignment");
        p.bias = nullptr; p.D = det_work; p.ldd = N; p.accumulate = 0; p.split_stride = (long long)M * N;
    }
    p.pool_n = 0; p.pool_row0 = 0; p.pool_idx0 = 0; p.pool_mask = nullptr; p.pool_max_u = nullptr; p.pool_max_m = nullptr;
    if (pool != nullptr) { p.pool_n = pool->n; p.pool_row0 = pool->row0; p.pool_idx0 = pool->idx0; p.pool_mask = pool->mask; p.pool_max_u = pool->max_u; p.pool_max_m = pool->max_m; }
    // diagnosis only (results are then wrong): WF_B200_POOL_NOPUSH=1 drops the RED.MAX.64 pushes of whole-cloud tiles, to separate
    // the cost of the column walk from the cost of the atomics
    { const char* e = getenv("WF_B200_POOL_NOPUSH"); if (pool != nullptr && e && e[0] == '1') p.pool_max_u = nullptr; }
    // bf16 tiles that are simply stored leave through TMA (WF_B200_GEMM_TMA_STORE=0 keeps the per-thread stores)
    static const bool tma_store_env = [] { const char* e = getenv("WF_B200_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
    p.tma_store = (tma_store_env && out_dtype == WF_BF16 && !p.accumulate && pool == nullptr && p.split_stride == 0) ? 1 : 0;
    CUtensorMap md = ma;                                         // placeholder when unused
    if (p.tma_store && (rc = make_store_map(&md, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd)) != WF_OK) return rc;
    // L2 eviction priorities (WF_B200_L2_HINTS=1; default: all normal).  Tried: evict_last on operands that several
    // concurrent tiles re-read, evict_first on the output and on once-read operands.  Measured (ncu, profiles/r02_gemm_traffic_*):
    // DRAM reads went UP (weight gradients 5.5 -> 6.8 GB, the 2048-wide layer 3.8 -> 4.1 GB) -- pinning everything thrashes -- so
    // the hints stay off.
    static const bool hints = [] { const char* e = getenv("WF_B200_L2_HINTS"); return e && e[0] == '1'; }();
    p.hint_a = p.hint_b = p.hint_d = ptx::L2_EVICT_NORMAL;
    if (hints) {
        p.hint_a = p.tiles_n > 1 ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_FIRST;           // A tile: read by every N tile of its rows
        p.hint_b = m_units > 1 ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_FIRST;             // B tile: read by every M unit
        p.hint_d = ptx::L2_EVICT_FIRST;
    }
    p.n_stages = 6; p.n_side = 0;
    for (int i = 0; i < n_segs; ++i) {
        const wf_side_seg& sg = segs[i];
        if (sg.rows <= 0) continue;
        const bool fwd = sg.kind == WF_SIDE_LN_FWD || sg.kind == WF_SIDE_LN_FWD_COLSUM;
        WF_CHECK_ARG(fwd || sg.kind == WF_SIDE_LN_BWD, "wf_gemm_bf16_side: segment %d: unknown kind %d", i, sg.kind);
        WF_CHECK_ARG(sg.kind == WF_SIDE_LN_FWD_COLSUM ? sg.C == 1024 : (sg.C == 1024 || sg.C == 2048),
                     "wf_gemm_bf16_side: segment %d: C=%d not built", i, sg.C);
        WF_CHECK_ARG(sg.rows < (1ll << 31), "wf_gemm_bf16_side: segment %d: too many rows", i);
        WF_CHECK_ARG(sg.x0 && sg.mean && sg.rstd && sg.gamma && sg.beta && sg.out && (fwd || (sg.x1 && sg.acc0 && sg.acc1 && sg.acc2)),
                     "wf_gemm_bf16_side: segment %d: null pointer", i);
        WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(sg.x0) | reinterpret_cast<uintptr_t>(sg.x1) | reinterpret_cast<uintptr_t>(sg.out) |
                       reinterpret_cast<uintptr_t>(sg.gamma) | reinterpret_cast<uintptr_t>(sg.beta)) & 15) == 0,
                     "wf_gemm_bf16_side: segment %d: 16-byte alignment required", i);
        if (sg.kind == WF_SIDE_LN_FWD_COLSUM)
            WF_CHECK_ARG(sg.part && sg.pool_n >= lnb::CS_R && sg.row_off >= 0 && sg.row_off % lnb::CS_R == 0,
                         "wf_gemm_bf16_side: segment %d: colsum needs part, points_per_cloud >= %d, row_off %% %d == 0", i, lnb::CS_R, lnb::CS_R);
        const int need = fwd ? (sg.kind == WF_SIDE_LN_FWD_COLSUM ? SIDE_SMEM_FWD_COLSUM_1024 : 0)
                             : (sg.C == 1024 ? SIDE_SMEM_BWD_1024 : SIDE_SMEM_BWD_2048);
        const int st = need > 32768 ? 4 : 5;
        if (st < p.n_stages) p.n_stages = st;
        p.side[p.n_side++] = sg;
    }
    p.own_kind = 0; p.own_gamma = p.own_beta = nullptr; p.own_h = nullptr; p.own_mean = p.own_rstd = nullptr; p.own_eps = 0.f; p.own_done = nullptr;
    if (own != nullptr) {
        WF_CHECK_ARG(mode == 2 && esz == 2 && a_kmajor && b_kmajor && p.tma_store && rowstats != nullptr && p.split_k == 1 &&
                     (N == 1024 || N == 2048) && N % BN == 0,
                     "wf_gemm_bf16_ownln: needs the 2-SM bf16 K-major form with row statistics, M >= 256, N in {1024, 2048}");
        WF_CHECK_ARG(own->gamma && own->beta && own->h && own->mean && own->rstd && own->done, "wf_gemm_bf16_ownln: null pointer");
        WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(own->h) | reinterpret_cast<uintptr_t>(own->gamma) | reinterpret_cast<uintptr_t>(own->beta)) & 15) == 0
                     && ldd == N, "wf_gemm_bf16_ownln: 16-byte alignment and a dense D (ldd == N) required");
        p.own_kind = 1; p.own_gamma = own->gamma; p.own_beta = own->beta; p.own_h = own->h; p.own_mean = own->mean; p.own_rstd = own->rstd;
        p.own_eps = own->eps; p.own_done = own->done;
    }
    if ((p.n_side > 0 || p.own_kind != 0) && p.n_stages == 6) p.n_stages = 5;
    long long items = p.streamk ? (long long)m_units * p.tiles_n * p.nkb          // units: any worker count up to this
                                : (long long)m_units * p.tiles_n * p.split_k;
    int workers = (int)(items < workers_max ? items : workers_max);
    if (group_workers > 0 && group_workers < workers) workers = group_workers;
    p.n_workers = workers;
    const int grid = mc ? 2 * workers : workers;
    const int key = (esz == 4 ? 4 : 0) | (a_kmajor ? 2 : 0) | (b_kmajor ? 1 : 0);
    if (p.n_side > 0 || p.own_kind != 0) {
        // every CTA of the launch takes its share of the side rows: always the full grid, whatever the GEMM's item count
        const int sgrid = 2 * workers_max;
        rc = a_kmajor ? launch<2, true, true, 2, true>(ma, mb, md, p, sgrid, stream) : launch<2, false, false, 2, true>(ma, mb, md, p, sgrid, stream);
        if (rc != WF_OK || !det) return rc;
    } else
    switch (key) {
        case 3: rc = launch_mc<2, true, true>(mode, ma, mb, md, p, grid, stream); break;
        case 2: rc = launch_mc<2, true, false>(mode, ma, mb, md, p, grid, stream); break;
        case 1: rc = launch_mc<2, false, true>(mode, ma, mb, md, p, grid, stream); break;
        case 0: rc = launch_mc<2, false, false>(mode, ma, mb, md, p, grid, stream); break;
        case 7: rc = launch_mc<4, true, true>(mode, ma, mb, md, p, grid, stream); break;
        case 6: rc = launch_mc<4, true, false>(mode, ma, mb, md, p, grid, stream); break;
        case 5: rc = launch_mc<4, false, true>(mode, ma, mb, md, p, grid, stream); break;
        default: rc = launch_mc<4, false, false>(mode, ma, mb, md, p, grid, stream); break;
    }
    if (rc != WF_OK || !det) return rc;
    const long long mn4 = (long long)M * N / 4;
    const int rgrid = (int)(cdiv(mn4, 256) < 8LL * sm_count() ? cdiv(mn4, 256) : 8LL * sm_count());
    splitk_reduce_kernel<<<rgrid, 256, 0, stream>>>(reinterpret_cast<const float4*>(det_work), p.split_k, mn4, N / 4,
                                                    reinterpret_cast<const float4*>(bias), static_cast<float*>(D), ldd, accumulate);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}}  // namespace wf::tc

extern "C" int wf_gemm_bf16(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, int M, int N,
                            int K, const float* bias, void* D, int ldd, int out_dtype, int accumulate, int split_k,
                            float* rowstats, wf_stream_t stream) {
    return wf::tc::gemm_tc(2, A, lda, a_kmajor, B, ldb, b_kmajor, M, N, K, bias, D, ldd, out_dtype, accumulate, split_k,
                           rowstats, wf::as_stream(stream));
}

extern "C" int wf_gemm_rowstats_parts(int N) { return 2 * wf::cdiv(N, wf::tc::BN); }

extern "C" int wf_gemm_bf16_pool(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                                 int points_per_cloud, int row_offset, int index_offset, const uint8_t* mask,
                                 uint64_t* max_u, uint64_t* max_m, wf_stream_t stream) {
    WF_CHECK_ARG(points_per_cloud >= 32 && row_offset >= 0 && index_offset >= 0,
                 "wf_gemm_bf16_pool: points_per_cloud must be >= 32 (got %d), row_offset/index_offset >= 0", points_per_cloud);
    WF_CHECK_ARG(max_u != nullptr && max_m != nullptr, "wf_gemm_bf16_pool: packed outputs required");
    wf::tc::PoolArgs pa{points_per_cloud, row_offset, index_offset, mask, reinterpret_cast<unsigned long long*>(max_u),
                        reinterpret_cast<unsigned long long*>(max_m)};
    return wf::tc::gemm_tc(2, A, lda, 1, B, ldb, 1, M, N, K, bias, nullptr, 8, WF_BF16, 0, 1, nullptr, wf::as_stream(stream), &pa);
}

extern "C" int wf_gemm_bf16_side(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, int M, int N, int K,
                                 const float* bias, void* D, int ldd, int out_dtype, int accumulate, int split_k, float* rowstats,
                                 int pool_n, int pool_row_offset, int pool_index_offset, const uint8_t* pool_mask,
                                 uint64_t* pool_max_u, uint64_t* pool_max_m, const wf_side_seg* segs, int n_segs,
                                 wf_stream_t stream) {
    WF_CHECK_ARG(n_segs >= 0 && (n_segs == 0 || segs != nullptr), "wf_gemm_bf16_side: bad segment list");
    if (pool_n > 0) {
        WF_CHECK_ARG(pool_n >= 32 && pool_row_offset >= 0 && pool_index_offset >= 0 && pool_max_u && pool_max_m && a_kmajor && b_kmajor,
                     "wf_gemm_bf16_side: pooling epilogue needs points_per_cloud >= 32, packed outputs and K-major operands");
        wf::tc::PoolArgs pa{pool_n, pool_row_offset, pool_index_offset, pool_mask, reinterpret_cast<unsigned long long*>(pool_max_u),
                            reinterpret_cast<unsigned long long*>(pool_max_m)};
        return wf::tc::gemm_tc(2, A, lda, 1, B, ldb, 1, M, N, K, bias, nullptr, 8, WF_BF16, 0, 1, nullptr, wf::as_stream(stream), &pa,
                               nullptr, 0, segs, n_segs);
    }
    return wf::tc::gemm_tc(2, A, lda, a_kmajor, B, ldb, b_kmajor, M, N, K, bias, D, ldd, out_dtype, accumulate, split_k, rowstats,
                           wf::as_stream(stream), nullptr, nullptr, 0, segs, n_segs);
}

extern "C" int wf_gemm_bf16_ownln(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, void* Z,
                                  float* rowstats, const float* gamma, const float* beta, void* H, float* mean, float* rstd,
                                  float eps, int32_t* done, const wf_side_seg* segs, int n_segs, wf_stream_t stream) {
    WF_CHECK_ARG(n_segs >= 0 && (n_segs == 0 || segs != nullptr), "wf_gemm_bf16_ownln: bad segment list");
    wf::tc::OwnArgs oa{gamma, beta, H, mean, rstd, eps, done};
    return wf::tc::gemm_tc(2, A, lda, 1, B, ldb, 1, M, N, K, bias, Z, N, WF_BF16, 0, 1, rowstats, wf::as_stream(stream), nullptr,
                           nullptr, 0, segs, n_segs, &oa);
}

extern "C" int wf_gemm_tf32(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, int M, int N,
                            int K, const float* bias, float* D, int ldd, int accumulate, int split_k, wf_stream_t stream) {
    return wf::tc::gemm_tc(4, A, lda, a_kmajor, B, ldb, b_kmajor, M, N, K, bias, D, ldd, WF_F32, accumulate, split_k,
                           nullptr, wf::as_stream(stream));
}

extern "C" int wf_gemm_tf32_splitk(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, int M, int N,
                                   int K, const float* bias, float* D, int ldd, int accumulate, int split_k, float* work,
                                   int64_t work_floats, wf_stream_t stream) {
    WF_CHECK_ARG(work != nullptr && work_floats >= 0, "wf_gemm_tf32_splitk: workspace required");
    return wf::tc::gemm_tc(4, A, lda, a_kmajor, B, ldb, b_kmajor, M, N, K, bias, D, ldd, WF_F32, accumulate, split_k,
                           nullptr, wf::as_stream(stream), nullptr, work, work_floats);
}
