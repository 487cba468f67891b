// Kernel 2: shared per-point MLP 8 -> 64 -> 128 -> 1024 + max over points on the 5th-generation
// tensor cores (tcgen05.mma, bf16 operands, fp32 accumulators in TMEM), sm_100a only.
// Reference call: model({"point_x": ...}), python/ossid/utils/zephyr_utils.py:34.
//
// Work split.  Persistent kernel, one CTA per SM, launched as clusters of two CTAs (the two SMs of a TPC) that
// execute every MMA together (cta_group::2, M = 256): CTA `rank` supplies 128 rows of the M side and half of the
// N side of each instruction and receives its own 128 accumulator lanes.  A pair walks the same hypotheses.
//
// Pair-tile = 256 consecutive points of one hypothesis; CTA `rank` owns points [128*rank, 128*rank+128) of it.
//   L1  D1[own 128 pts x  64] = X [256 x 16(8 real + 8 zero)] . W1^T     M=256 N=64  K=16  (1 MMA;  W1 rows split 32/32)
//   L2  D2[own 128 pts x 128] = H1[256 x 64]  . W2^T                     M=256 N=128 K=64  (4 MMAs; W2 rows split 64/64)
//   L3  D3[own 128 ch  x 128 pts] = W3blk[256 ch x 128] . H2sub^T        M=256 N=128 K=128 (8 MMAs) for each of
//       4 channel blocks x 2 point halves: the M side is the pair's 2 x 128 channels of block cb (CTA `rank` keeps
//       channels [512*rank, 512*rank+512) of W3, 128 KB bf16, resident in its shared memory for the whole launch),
//       the N side is points [64*h, 64*h+64) of EACH CTA's H2 tile.
// So layers 1-2 are computed once per point (the single-CTA version recomputed them in both CTAs), every L3
// instruction reads 4 KB of A + 2 KB of B per CTA instead of 4 + 4 (N = 128 single-CTA MMAs need all 128 B/clk of
// shared-memory bandwidth), and the front epilogue handles half as many rows per point of work -- the kernel is
// power-bound in sustained runs, so the energy per hypothesis is what sets the throughput.
// L1/L2 put points on TMEM lanes, so the epilogue thread of a point holds its whole channel row and
// writes it as the next layer's K-major, 128B-swizzled operand with 16-byte stores.  L3 puts
// CHANNELS on lanes and points on columns, so max-over-points is a register-only reduction in the
// thread that owns the channel; bias + ReLU commute with max and are applied once per hypothesis.
//
// Warp roles (320 threads per CTA): warps 0-3 front epilogues (D1 -> H1, D2 -> H2), warps 4-7 max-pool
// epilogue (D3), warp 8 bulk-copy producer (weights once, then feature tiles, 3 stages), warp 9
// TMEM allocator and -- in the leader CTA (rank 0) only -- the single elected MMA-issuing thread.  Completion
// of MMAs reaches both CTAs through multicast tcgen05.commit; "operand ready" / "accumulator drained" travel
// the other way as per-warp mbarrier arrivals on the leader's barriers (count 8 = 4 epilogue warps x 2 CTAs).
// There are no padding rows (see tile_row0), so the max-pool epilogue is a plain max over every accumulator column.
// Front of pair-tile i+1 overlaps layer 3 of pair-tile i; D3 is triple-buffered in TMEM
// (cols: D1 0-63 inside D2 0-127, D3 128-255 / 256-383 / 384-511).  H1 aliases the H2
// buffer that the same tile's epilogue 2 overwrites afterwards, which is what makes two H2
// buffers + the resident weights fit in 227 KB.
#include <cuda.h>

#include "zs_common.cuh"

namespace {

// Optional per-role wait accounting (build with -DZS_TC_PROF; tools/k2_prof.py reads the counters out of dbg_h2).
#ifdef ZS_TC_PROF
#include <cstdlib>
__device__ int g_exp = 0;          // experiment bits (ZS_TC_EXPERIMENT): 1 = no operand stores, 2 = no bias loads, 4 = no TMEM loads in max-pool
#define EXP(bit) (g_exp & (bit))
constexpr bool kDbgDump = false;
#define PROF_DECL long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64()
#define PROF_WAIT(slot, stmt) do { const long long t_ = clock64(); stmt; prof[slot] += clock64() - t_; } while (0)
#define PROF_DUMP(base, cnt) do { if (dbg_h2) { long long* o_ = reinterpret_cast<long long*>(dbg_h2) + (size_t)blockIdx.x * 24 + (base); \
        o_[0] = clock64() - prof_t0; for (int q_ = 0; q_ < (cnt); ++q_) o_[1 + q_] = prof[q_]; } } while (0)
#else
#define EXP(bit) false
constexpr bool kDbgDump = true;
#define PROF_DECL
#define PROF_WAIT(slot, stmt) stmt
#define PROF_DUMP(base, cnt)
#endif

}  // namespace
#include "zs_tc.cuh"     // after EXP(): the shared epilogue helpers honour the profiling build's experiment bits
namespace {

constexpr int kTile = 128;               // points per CTA per pair-tile
constexpr int kPairTile = 2 * kTile;     // points per pair-tile
constexpr int kThreadsTc = 320;
constexpr int kThreadsFused = 352;       // + warp 10: the second feature-producer warp of the fused kernel
constexpr int kStages = 3;

// Fused variant (K1 inside K2): instead of bulk-copying bf16 feature rows that zs_features wrote to HBM, two producer
// warps per CTA project / gather / featurise the 128 points of each half-tile themselves (the per-point code of
// zs_common.cuh, so the rows are bit-identical) and store them straight into the X stage.  The hypothesis list of a
// launch may span several objects of one cloud size: segment g covers rows [first_row[g], first_row[g+1]).
constexpr int kMaxFusedSegs = 32;
struct fused_seg {
    zs_obj_view o;            // model cloud (global memory; 36 B per point, read coalesced through L1 / L2)
    const float* poses;       // poses of this segment's object
    const int32_t* keep_idx;  // kept list of the pre-filter (nullable): hypothesis i of the segment is pose keep_idx[i]
    const int32_t* n_dev;     // device-side count of the segment (nullable = cap): zs_filter's n_keep_out
    int first_row, cap;       // first pooled row of the segment (rows are laid out by capacity) and its capacity
};
struct fused_args {
    zs_cam cam;
    const float4* frame;    // packed frame {depth, H, S, V}
    int n_seg, pad_;
    fused_seg seg[kMaxFusedSegs];
};

// ---- shared-memory map (bytes from a 1024-aligned base) ---------------------------------------
constexpr uint32_t kSmW3 = 0;                          // 4 blocks x 2 k-halves x 16 KB
constexpr uint32_t kSmW2 = kSmW3 + 131072;             // 8 KB: this CTA's 64 rows of W2
constexpr uint32_t kSmA3 = kSmW2 + 8192;               // 2 x 32 KB  (H2; first 16 KB doubles as H1)
constexpr uint32_t kSmW1 = kSmA3 + 2 * 32768;          // 1 KB: this CTA's 32 rows of W1 (2 KB reserved)
constexpr uint32_t kSmX = kSmW1 + 2048;                // 3 x 2 KB feature tiles
constexpr uint32_t kSmZero = kSmX + kStages * 2048;    // 2 KB of zeros (upper K half of X)
constexpr uint32_t kSmB1 = kSmZero + 2048;             // 64 floats
constexpr uint32_t kSmB2 = kSmB1 + 256;                // 128 floats
constexpr uint32_t kSmBar = kSmB2 + 512;               // mbarriers
constexpr uint32_t kSmTmemPtr = kSmBar + 256;
constexpr uint32_t kSmBytes = kSmTmemPtr + 16;
constexpr uint32_t kSmAlloc = kSmBytes + 1024;         // slack for manual 1024-B alignment
static_assert(kSmAlloc <= 232448, "exceeds 227 KB of shared memory per CTA");

// barrier slots
enum : int {
    BAR_W_FULL = 0, BAR_X_FULL = 1 /*3*/, BAR_X_EMPTY = 4 /*3*/, BAR_D1_FULL = 7, BAR_A2_FULL = 8, BAR_D2_FULL = 9,
    BAR_A3_FULL = 10 /*2*/, BAR_A3_EMPTY = 12 /*2*/, BAR_D3_FULL = 14 /*3*/, BAR_D3_EMPTY = 17 /*3*/,
    BAR_XP_FULL = 20 /*3: peer's feature tile landed (leader only)*/, BAR_WP_FULL = 23 /*peer's weights landed*/, BAR_COUNT = 24
};

// TMEM columns
// D1 (64 cols) lives inside D2's 128 columns: D1 is dead once epilogue 1 has signalled A2_FULL, which the
// issuer waits for before layer 2 overwrites the range, and layer 1 of the next tile is only issued after
// A3_FULL of this tile (epilogue 2 has drained D2).  That leaves room for THREE layer-3 accumulators.
constexpr uint32_t kColD1 = 0, kColD2 = 0, kColD3 = 128;
constexpr int kD3Bufs = 3;
#ifndef ZS_L2_AFTER
#define ZS_L2_AFTER 1
#endif
constexpr int kL2After = ZS_L2_AFTER;     // layer 2 of pair-tile i is issued behind this many channel blocks of layer 3 of pair-tile i-1
constexpr uint32_t kTmemCols = 512;

// global image of the bf16 operands (bytes): W3 half 0, W3 half 1, W2, W1
constexpr size_t kImgW3Half = 131072, kImgW2 = 2 * kImgW3Half, kImgW1 = kImgW2 + 16384, kImgBytes = kImgW1 + 2048;

using namespace zs_tc;

// ---- the kernel ---------------------------------------------------------------------------------
template <bool kFused>
__global__ void __launch_bounds__(kFused ? kThreadsFused : kThreadsTc, 1)
zs_k_mlp_tc(const __nv_bfloat16* __restrict__ feat, int n, int N, const uint8_t* __restrict__ wimg,
            const float* __restrict__ wf32, float* __restrict__ pooled, float* __restrict__ dbg_h1,
            float* __restrict__ dbg_h2, const int32_t* __restrict__ n_dev, int n_off,
            const __grid_constant__ fused_args fa) {
    constexpr int kThreadsK = kFused ? kThreadsFused : kThreadsTc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Fused: the hypothesis list is the concatenation of the segments' (possibly device-side) counts.  Every warp keeps
    // the dense first index of segment `lane` in a register: "which segment owns hypothesis h" is one ballot.
    int seg_first = 0x7fffffff;
    if (kFused) {
        int cnt = 0;
        if (lane < fa.n_seg) {
            const int32_t* nd = fa.seg[lane].n_dev;
            cnt = nd ? min(fa.seg[lane].cap, max(0, __ldg(nd))) : fa.seg[lane].cap;
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        n = __shfl_sync(0xffffffffu, incl, 31);
        if (lane < fa.n_seg) seg_first = incl - cnt;
    } else {
        n = zs_dyn_count(n_dev, n_off, n);      // zs_set_dynamic_count: the count may live on the device
    }
    // (warp-uniform h) -> segment index and index inside the segment
    auto seg_of = [&](int h, int& local) {
        const int g = __popc(__ballot_sync(0xffffffffu, seg_first <= h)) - 1;
        local = h - __shfl_sync(0xffffffffu, seg_first, g);
        return g;
    };
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - smem_u32(smem_raw));
    auto bar = [&](int i) { return sbase + kSmBar + 8u * (uint32_t)i; };

    const int rank = (int)cluster_rank();                 // 0 = leader (issues every MMA of the pair)
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int T = (N + kPairTile - 1) / kPairTile;        // pair-tiles per hypothesis
    const int n_loc = pair < n ? (n - pair + n_pairs - 1) / n_pairs : 0;
    const int total = n_loc * T;

    // ---- one-time setup ------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(bar(BAR_W_FULL), 1);
        mbar_init(bar(BAR_WP_FULL), 1);
        for (int s = 0; s < kStages; ++s) {
            // fused: the leader's X_FULL collects one arrival per producer warp of the pair (2 + 2); XP_FULL is unused
            mbar_init(bar(BAR_X_FULL + s), kFused ? 4 : 1); mbar_init(bar(BAR_X_EMPTY + s), 1); mbar_init(bar(BAR_XP_FULL + s), 1);
        }
        mbar_init(bar(BAR_D1_FULL), 1); mbar_init(bar(BAR_D2_FULL), 1);
        mbar_init(bar(BAR_A2_FULL), 8);                                   // one arrival per epilogue warp of the pair (4 + 4)
        for (int b = 0; b < 2; ++b) { mbar_init(bar(BAR_A3_FULL + b), 8); mbar_init(bar(BAR_A3_EMPTY + b), 1); }
        for (int b = 0; b < kD3Bufs; ++b) { mbar_init(bar(BAR_D3_FULL + b), 1); mbar_init(bar(BAR_D3_EMPTY + b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the feature stages and the zero block (stale bytes must be finite), stage the small biases
    for (int i = tid; i < (int)((kStages + 1) * 2048 / 16); i += kThreadsK)
        reinterpret_cast<uint4*>(sm + kSmX)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 64; i += kThreadsK) reinterpret_cast<float*>(sm + kSmB1)[i] = wf32[ZS_OFF_B1 + i];
    for (int i = tid; i < 128; i += kThreadsK) reinterpret_cast<float*>(sm + kSmB2)[i] = wf32[ZS_OFF_B2 + i];
    fence_proxy_async();
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmTmemPtr), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                       // both CTAs: barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kSmTmemPtr);

    // No padding rows: a half-tile that would run past the hypothesis' last point is shifted back to end at it
    // (re-reading points the pair already has -- max-pooling ignores duplicates), and a hypothesis with fewer than
    // 128 points fills the tile with repeats of itself.  Every accumulator column is then a real point of the right
    // hypothesis and the max-pool epilogue needs no masking.
    // First feature row of this CTA's half of pair-tile `tt` of local hypothesis `j` (N >= kTile):
    auto tile_row0 = [&](int j, int tt) {
        const int s0 = tt * kPairTile + rank * kTile;
        return (long long)(pair + j * n_pairs) * N + (long long)(s0 + kTile <= N ? s0 : (N >= kTile ? N - kTile : 0));
    };

    if (kFused && (warp == 8 || warp == 10)) {
        // ===== feature producers (fused): warp 8 -> rows [0,64) of the half-tile, warp 10 -> rows [64,128) ==========
        if (warp == 8 && lane == 0) {                           // the weights still arrive by bulk copy, once
            mbar_expect_tx(bar(BAR_W_FULL), 131072 + 8192 + 1024);
            for (int c = 0; c < 8; ++c)
                bulk_g2s(sbase + kSmW3 + c * 16384, wimg + (size_t)rank * kImgW3Half + (size_t)c * 16384, 16384, bar(BAR_W_FULL));
            bulk_g2s(sbase + kSmW2, wimg + kImgW2 + (size_t)rank * 8192, 8192, bar(BAR_W_FULL));
            bulk_g2s(sbase + kSmW1, wimg + kImgW1 + (size_t)rank * 1024, 1024, bar(BAR_W_FULL));
            if (rank != 0) { mbar_wait(bar(BAR_W_FULL), 0); mbar_arrive_leader(leader_addr(bar(BAR_WP_FULL))); }
        }
        __syncwarp();
        const int row0 = warp == 8 ? 0 : 64;
        const uint32_t x_full = leader_addr(bar(BAR_X_FULL));
        const float4* __restrict__ frame = fa.frame;
        zs_pose P;
        const float4 *pA = nullptr, *pB = nullptr;
        const float* pV = nullptr;
        for (int i = 0; i < total; ++i) {
            const int s = i % kStages, j = i / T, tt = i - j * T;
            if (tt == 0) {                                      // new hypothesis: its object (segment) and pose
                int local;
                const int g = seg_of(pair + j * n_pairs, local);
                pA = fa.seg[g].o.pA; pB = fa.seg[g].o.pB; pV = fa.seg[g].o.pV;
                const int32_t* keep = fa.seg[g].keep_idx;
                P = zs_load_pose(fa.seg[g].poses, keep ? __ldg(keep + local) : local);
            }
            const int s0 = tt * kPairTile + rank * kTile;
            const int p0 = (s0 + kTile <= N ? s0 : N - kTile) + row0;      // first of this warp's 64 points (N >= kTile)
            float4 a[2], px[2];
            float x[2], y[2], z[2], uf[2], vf[2];
            bool valid[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {                    // phase 1: exact projection, issue the frame gather
                int pix;
                a[jj] = __ldg(pA + p0 + jj * 32 + lane);
                zs_feat_phase1(P, fa.cam, a[jj], x[jj], y[jj], z[jj], uf[jj], vf[jj], valid[jj], pix);
                px[jj] = __ldg(frame + pix);
            }
            uint4 v[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {                    // phase 2: residual features -> the 16-byte bf16 row
                const float4 b = __ldg(pB + p0 + jj * 32 + lane);
                const float vm = __ldg(pV + p0 + jj * 32 + lane);
                float f[7];
                zs_feat_phase2(P, fa.cam, a[jj], b, vm, px[jj], x[jj], y[jj], z[jj], uf[jj], vf[jj], f);
                v[jj] = zs_feat_row_bf16(f, valid[jj]);
            }
            mbar_wait(bar(BAR_X_EMPTY + s), ((i / kStages) & 1) ^ 1);      // layer 1 of pair-tile i-3 has read this stage
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
                *reinterpret_cast<uint4*>(sm + kSmX + s * 2048 + (row0 + jj * 32 + lane) * 16) = v[jj];
            fence_proxy_async();                                // generic-proxy stores -> visible to the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(x_full + 8u * s);
        }
    } else if (!kFused && warp == 8) {
        // ===== bulk-copy producer =================================================================
        if (lane == 0) {
            mbar_expect_tx(bar(BAR_W_FULL), 131072 + 8192 + 1024);
            for (int c = 0; c < 8; ++c)
                bulk_g2s(sbase + kSmW3 + c * 16384, wimg + (size_t)rank * kImgW3Half + (size_t)c * 16384, 16384, bar(BAR_W_FULL));
            bulk_g2s(sbase + kSmW2, wimg + kImgW2 + (size_t)rank * 8192, 8192, bar(BAR_W_FULL));
            bulk_g2s(sbase + kSmW1, wimg + kImgW1 + (size_t)rank * 1024, 1024, bar(BAR_W_FULL));
            // The leader's issuer must also know that the PEER's operands have landed: the peer's producer
            // relays its own completions (one tile behind its copies) to barriers in the leader's shared memory.
            const bool relay = rank != 0;
            const uint32_t wp = leader_addr(bar(BAR_WP_FULL)), xp = leader_addr(bar(BAR_XP_FULL));
            if (relay) { mbar_wait(bar(BAR_W_FULL), 0); mbar_arrive_leader(wp); }
            for (int i = 0; i < total; ++i) {
                const int s = i % kStages, j = i / T, tt = i - j * T;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(feat) + tile_row0(j, tt) * 16;
                mbar_wait(bar(BAR_X_EMPTY + s), ((i / kStages) & 1) ^ 1);
                mbar_expect_tx(bar(BAR_X_FULL + s), 2048);
                if (N >= kTile) {
                    bulk_g2s(sbase + kSmX + s * 2048, src, 2048, bar(BAR_X_FULL + s));
                } else {
                    for (int r = 0; r < kTile; r += N)      // the whole (short) hypothesis, repeated
                        bulk_g2s(sbase + kSmX + s * 2048 + r * 16, src, (uint32_t)min(N, kTile - r) * 16, bar(BAR_X_FULL + s));
                }
                if (relay && i >= 1) {
                    const int ip = i - 1, sp = ip % kStages;
                    mbar_wait(bar(BAR_X_FULL + sp), (ip / kStages) & 1);
                    mbar_arrive_leader(xp + 8u * sp);
                }
            }
            if (relay && total >= 1) {
                const int ip = total - 1, sp = ip % kStages;
                mbar_wait(bar(BAR_X_FULL + sp), (ip / kStages) & 1);
                mbar_arrive_leader(xp + 8u * sp);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer (one thread of the leader CTA) ==========================================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc_l1 = make_idesc(256, 64), idesc_128 = make_idesc(256, 128);
            mbar_wait(bar(BAR_W_FULL), 0);
            mbar_wait(bar(BAR_WP_FULL), 0);
            tc_fence_after();
            PROF_DECL;
            // Descriptors differ only in their 14-bit start-address field, so every MMA's pair is "base low word +
            // immediate": the issuing thread must not spend more than ~64 cycles of dependent address arithmetic
            // per MMA or it, not the tensor pipe, paces the kernel (tools/mma_probe.cu measures exactly that).
            // A descriptor names the same offset in both CTAs' shared memory.
            constexpr uint32_t hi_sw = desc_hi(1024, kLayoutSw128), hi_x = desc_hi(128, kLayoutNone);
            const uint32_t w3_lo = desc_lo(sbase + kSmW3, 16), w2_lo = desc_lo(sbase + kSmW2, 16);
            const uint32_t w1_lo = desc_lo(sbase + kSmW1, 512);
            auto issue_l3 = [&](int it, int cb, int hh) {   // layer 3 of pair-tile `it`: channel block cb, point half hh
                const int q = it * 8 + cb * 2 + hh, b = q % kD3Bufs, buf = it & 1;   // A3_FULL of `it` was awaited by the caller
                const uint32_t a3_lo = desc_lo(sbase + kSmA3 + buf * 32768 + hh * 8192, 16), d3 = tmem + kColD3 + b * 128;
                PROF_WAIT(1, mbar_wait(bar(BAR_D3_EMPTY + b), ((q / kD3Bufs) & 1) ^ 1));
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t kb = k >> 2, kk = k & 3;
                    tc_mma(d3, desc_at(w3_lo, hi_sw, (cb * 2 + kb) * 16384 + kk * 32),
                           desc_at(a3_lo, hi_sw, kb * 16384 + kk * 32), idesc_128, k > 0);
                }
                tc_commit(bar(BAR_D3_FULL + b));
                if (cb == 3 && hh == 1) tc_commit(bar(BAR_A3_EMPTY + buf));
            };
            for (int i = 0; i <= total; ++i) {
                // H2 of pair-tile i-1 is ready in both CTAs and D2 (which D1 aliases) has been drained by both epilogues 2
                if (i >= 1) PROF_WAIT(0, mbar_wait(bar(BAR_A3_FULL + ((i - 1) & 1)), ((i - 1) >> 1) & 1));
                if (i < total) {
                    const int s = i % kStages;
                    PROF_WAIT(3, mbar_wait(bar(BAR_X_FULL + s), (i / kStages) & 1));
                    if (!kFused) PROF_WAIT(3, mbar_wait(bar(BAR_XP_FULL + s), (i / kStages) & 1));
                    tc_fence_after();
                    const uint32_t xa = sbase + kSmX + s * 2048;
                    tc_mma(tmem + kColD1, desc_at(desc_lo(xa, (sbase + kSmZero) - xa), hi_x, 0), desc_at(w1_lo, hi_x, 0), idesc_l1, 0);
                    tc_commit(bar(BAR_X_EMPTY + s));
                    tc_commit(bar(BAR_D1_FULL));
                }
                if (i >= 1) { issue_l3(i - 1, 0, 0); issue_l3(i - 1, 0, 1); if (kL2After >= 2) { issue_l3(i - 1, 1, 0); issue_l3(i - 1, 1, 1); } }
                if (i < total) {
                    const uint32_t a2_lo = desc_lo(sbase + kSmA3 + (i & 1) * 32768, 16);
                    PROF_WAIT(2, mbar_wait(bar(BAR_A2_FULL), i & 1));
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma(tmem + kColD2, desc_at(a2_lo, hi_sw, kk * 32), desc_at(w2_lo, hi_sw, kk * 32), idesc_128, kk > 0);
                    tc_commit(bar(BAR_D2_FULL));
                }
                if (i >= 1) {
#pragma unroll
                    for (int cb = kL2After; cb < 4; ++cb) { issue_l3(i - 1, cb, 0); issue_l3(i - 1, cb, 1); }
                }
            }
            PROF_DUMP(0, 4);
        }
    } else if (warp < 4) {
        // ===== front epilogues: D1 -> H1 (bf16, swizzled), D2 -> H2 ================================
        const uint32_t r = (uint32_t)tid;                               // TMEM lane = point row of this CTA's half-tile
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const float* b1 = reinterpret_cast<const float*>(sm + kSmB1);
        const float* b2 = reinterpret_cast<const float*>(sm + kSmB2);
        const uint32_t a2_full = leader_addr(bar(BAR_A2_FULL)), a3_full = leader_addr(bar(BAR_A3_FULL));
        PROF_DECL;
        for (int i = 0; i < total; ++i) {
            const int buf = i & 1, j = i / T, tt = i - j * T;
            const long long drow = tile_row0(j, tt) + (N >= kTile ? (int)r : (int)r % N);   // the point this row holds
            uint8_t* a3 = sm + kSmA3 + buf * 32768;
            PROF_WAIT(0, mbar_wait(bar(BAR_A3_EMPTY + buf), ((i >> 1) & 1) ^ 1));     // layer 3 of pair-tile i-2 has released this buffer
            PROF_WAIT(1, mbar_wait(bar(BAR_D1_FULL), i & 1));
            tc_fence_after();
            {   // layer 1: 64 channels = two 32-column loads in flight, one wait
                uint32_t v0[32], v1[32];
                tc_ld32(lane_addr + kColD1, v0);
                tc_ld32(lane_addr + kColD1 + 32, v1);
                tc_wait_ld();
                epi_store32(v0, b1, a3, r, 0);
                epi_store32(v1, b1 + 32, a3, r, 4);
                if (kDbgDump && dbg_h1) {
                    dbg_dump32(dbg_h1 + drow * 64, v0, b1);
                    dbg_dump32(dbg_h1 + drow * 64 + 32, v1, b1 + 32);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(a2_full);

            PROF_WAIT(2, mbar_wait(bar(BAR_D2_FULL), i & 1));
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {                            // layer 2: 2 x 64 channels (= K halves of layer 3)
                uint32_t v0[32], v1[32];
                tc_ld32(lane_addr + kColD2 + kb * 64, v0);
                tc_ld32(lane_addr + kColD2 + kb * 64 + 32, v1);
                tc_wait_ld();
                epi_store32(v0, b2 + kb * 64, a3 + kb * 16384, r, 0);
                epi_store32(v1, b2 + kb * 64 + 32, a3 + kb * 16384, r, 4);
                if (kDbgDump && dbg_h2) {
                    dbg_dump32(dbg_h2 + drow * 128 + kb * 64, v0, b2 + kb * 64);
                    dbg_dump32(dbg_h2 + drow * 128 + kb * 64 + 32, v1, b2 + kb * 64 + 32);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(a3_full + 8u * buf);
        }
        if (tid == 0) PROF_DUMP(8, 3);
    } else if (warp < 8) {
        // ===== max-pool epilogue: D3[channel lane][point column] -> running max -> pooled ===========
        // Columns of a layer-3 accumulator of point half hh: 0-63 = points 64*hh.. of the leader's half-tile,
        // 64-127 = points 64*hh.. of the peer's half-tile (the N side concatenates the two CTAs' operand rows).
        // Every column is a real point (see tile_row0), so this is a plain max.  The point-half loop is kept
        // rolled: eight unrolled copies of the body overflow the instruction cache.
        const int wq = warp - 4;                                        // TMEM lane quadrant == warp % 4
        const int L = wq * 32 + lane;                                   // channel within the 128-channel block
        const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
        const uint32_t d3_empty = leader_addr(bar(BAR_D3_EMPTY));
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        PROF_DECL;
        for (int i = 0; i < total; ++i) {
            const int j = i / T, tt = i - j * T;
            int h = pair + j * n_pairs;                                 // pooled row of the hypothesis
            if (kFused && tt == T - 1) {                                // rows are laid out by segment capacity
                int local;
                const int g = seg_of(h, local);
                h = fa.seg[g].first_row + local;
            }
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float qq[4] = {m[cb], -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const int q = i * 8 + cb * 2 + hh, b = q % kD3Bufs;
                    PROF_WAIT(0, mbar_wait(bar(BAR_D3_FULL + b), (q / kD3Bufs) & 1));
                    tc_fence_after();
                    {
                        uint32_t v0[32], v1[32], v2[32], v3[32];           // all 128 point columns in flight, one wait
                        const uint32_t a = lane_addr + kColD3 + b * 128;
                        PROF_WAIT(1, if (!EXP(4)) { tc_ld32(a, v0); tc_ld32(a + 32, v1); tc_ld32(a + 64, v2); tc_ld32(a + 96, v3); }
                                     tc_wait_ld());
                        // the accumulator is free as soon as it sits in registers: release it before the arithmetic
                        tc_fence_before();
                        __syncwarp();
                        PROF_WAIT(3, if (lane == 0) mbar_arrive_leader(d3_empty + 8u * b));
                        max32(v0, qq); max32(v1, qq); max32(v2, qq); max32(v3, qq);
                    }
                }
                float mm = fmaxf(fmaxf(qq[0], qq[1]), fmaxf(qq[2], qq[3]));
                if (tt == T - 1) {
                    const int ch = rank * 512 + cb * 128 + L;
                    pooled[(size_t)h * 1024 + ch] = fmaxf(mm + __ldg(wf32 + ZS_OFF_B3 + ch), 0.f);
                    mm = -INFINITY;
                }
                m[cb] = mm;
            }
        }
        if (tid == 128) PROF_DUMP(16, 4);
    }

    // ---- teardown --------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync();                        // the peer may still be signalling this CTA's barriers / reading its operands
    if (warp == 9) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ---- weight images -------------------------------------------------------------------------------
// Writes the bf16 operand images exactly as they sit in shared memory (see the map above).
__global__ void zs_k_build_images(const float* __restrict__ w, uint8_t* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // W3: 1024 x 128
    if (i < 1024 * 128) {
        const int ch = i >> 7, k = i & 127;
        const int half = ch >> 9, cb = (ch >> 7) & 3, mrow = ch & 127, kb = k >> 6, kc = k & 63;
        const size_t off = (size_t)half * kImgW3Half + (size_t)(cb * 2 + kb) * 16384 + sw128_off(mrow, kc >> 3) + (kc & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(w[ZS_OFF_W3 + i]);
    }
    // W2: 128 x 64; rows 0-63 (leader's share of the N side) and 64-127 (peer's) are the two 8 KB halves of the image
    if (i < 128 * 64) {
        const int ch = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(img + kImgW2 + sw128_off(ch, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(w[ZS_OFF_W2 + i]);
    }
    // W1: 64 x 16 (K 8..15 zero), no swizzle, chunk-major core matrices; rows 0-31 (leader's share of the N side)
    // and rows 32-63 (peer's) are separate 1 KB images
    if (i < 64 * 16) {
        const int ch = i >> 4, k = i & 15, chl = ch & 31;
        const size_t off = (size_t)(ch >> 5) * 1024 + (size_t)(k >> 3) * 512 + (size_t)(chl >> 3) * 128 + (chl & 7) * 16 + (k & 7) * 2;
        const float v = k < 8 ? w[ZS_OFF_W1 + ch * 8 + k] : 0.f;
        *reinterpret_cast<__nv_bfloat16*>(img + kImgW1 + off) = __float2bfloat16_rn(v);
    }
}

}  // namespace

int zs_tc_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st) {
    zs_weights& w = ctx->w[slot];
    if (!w.bf16 && cudaMalloc(&w.bf16, kImgBytes) != cudaSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_NOMEM, "bf16 weight images");
    }
    zs_k_build_images<<<(1024 * 128 + 255) / 256, 256, 0, st>>>(w.f32, reinterpret_cast<uint8_t*>(w.bf16));
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

void zs_tc_destroy(zs_ctx*) {}

template <bool kFused>
static int launch_tc(zs_ctx* ctx, int slot, const __nv_bfloat16* feat, int n, int n_pts, float* pooled,
                     float* dbg_h1, float* dbg_h2, const fused_args& fa, cudaStream_t st) {
    const zs_weights& w = ctx->w[slot];
    ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_mlp_tc<kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmAlloc));
#ifdef ZS_TC_PROF
    { const char* e = getenv("ZS_TC_EXPERIMENT"); int v = e ? atoi(e) : 0; cudaMemcpyToSymbolAsync(g_exp, &v, sizeof(int), 0, cudaMemcpyHostToDevice, st); }
#endif
    int grid = ctx->sm_count & ~1;                 // CTA pairs
    if (grid > 2 * n) grid = 2 * n;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kFused ? kThreadsFused : kThreadsTc, 1, 1);
    cfg.dynamicSmemBytes = kSmAlloc;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;    // CTA pair = the two SMs of a TPC (cta_group::2 MMAs)
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ZS_CUDA(ctx, cudaLaunchKernelEx(&cfg, zs_k_mlp_tc<kFused>, feat, n, n_pts, reinterpret_cast<const uint8_t*>(w.bf16),
                                    (const float*)w.f32, pooled, dbg_h1, dbg_h2, kFused ? nullptr : ctx->dyn_n,
                                    kFused ? 0 : ctx->dyn_off, fa));
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

int zs_score_tc(zs_ctx* ctx, int slot, const __nv_bfloat16* feat, int n, int n_pts, float* pooled, cudaStream_t st) {
    return launch_tc<false>(ctx, slot, feat, n, n_pts, pooled, nullptr, nullptr, fused_args{}, st);
}

// Fused projection + gather + features + shared MLP + max-pool (bf16): the hypothesis list is the concatenation of the
// segments (segment i: n_hyp[i] hypotheses of the cloud in obj_slots[i], poses [dev] float32 [n_hyp[i]][12]).
extern "C" int zs_pool_fused(zs_ctx* ctx, int weight_slot, int n_seg, const int32_t* obj_slots, const float* const* poses,
                             const int32_t* const* keep_idx, const int32_t* n_hyp, const int32_t* const* n_dev,
                             float* pooled_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (weight_slot < 0 || weight_slot >= ZS_MAX_WEIGHT_SLOTS || !ctx->w[weight_slot].set)
        return zs_fail(ctx, ZS_ERR_STATE, "weight slot %d not set", weight_slot);
    if (n_seg < 0 || (n_seg > 0 && (!obj_slots || !poses || !n_hyp || !pooled_out)))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_pool_fused arguments");
    if (ctx->dyn_n) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "zs_set_dynamic_count is not used by zs_pool_fused: pass n_dev per segment");
    if (!ctx->frame.set) return zs_fail(ctx, ZS_ERR_STATE, "frame not set");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const zs_frame& f = ctx->frame;
    size_t row = 0;
    int i = 0, n_pts = 0;
    while (i < n_seg) {                                   // launches of at most kMaxFusedSegs segments
        fused_args fa = {};
        fa.cam = zs_cam{f.fx, f.fy, f.cx, f.cy, f.inv_fx, f.inv_fy, f.H, f.W};
        fa.frame = f.packed;
        int n = 0;
        for (; i < n_seg && fa.n_seg < kMaxFusedSegs; ++i) {
            if (n_hyp[i] == 0) continue;
            const int slot = obj_slots[i];
            if (slot < 0 || slot >= ZS_MAX_OBJECTS || ctx->obj[slot].n_pts == 0)
                return zs_fail(ctx, ZS_ERR_STATE, "object slot %d not set", slot);
            const zs_object& ob = ctx->obj[slot];
            if (n_pts == 0) n_pts = ob.n_pts;
            if (ob.n_pts != n_pts || n_pts < kTile)
                return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "fused scoring needs one cloud size >= %d per call (got %d and %d)",
                               kTile, n_pts, ob.n_pts);
            if (n_hyp[i] < 0 || !poses[i] || ((uintptr_t)poses[i] & 15))
                return zs_fail(ctx, ZS_ERR_INVALID, "segment %d: n_hyp %d / poses alignment", i, n_hyp[i]);
            fused_seg& g = fa.seg[fa.n_seg++];
            g.o = zs_obj_view{ob.pA, ob.pB, ob.pV, ob.n_pts};
            g.poses = poses[i];
            g.keep_idx = keep_idx ? keep_idx[i] : nullptr;
            g.n_dev = n_dev ? n_dev[i] : nullptr;
            g.first_row = n;
            g.cap = n_hyp[i];
            n += n_hyp[i];
        }
        if (n == 0) continue;
        int rc = launch_tc<true>(ctx, weight_slot, nullptr, n, n_pts, pooled_out + row * 1024, nullptr, nullptr, fa,
                                 (cudaStream_t)stream);
        if (rc) return rc;
        row += (size_t)n;
    }
    return ZS_OK;
}

// Diagnostic entry point: as zs_pool (bf16 or split-bf16 features) but also dumps the activations of layers 1 and 2 as
// the next layer reads them (bf16-rounded; hi + lo for the split kernel).
extern "C" int zs_pool_debug(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts, float* pooled_out,
                             float* h1_out, float* h2_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (weight_slot < 0 || weight_slot >= ZS_MAX_WEIGHT_SLOTS || !ctx->w[weight_slot].set)
        return zs_fail(ctx, ZS_ERR_STATE, "weight slot %d not set", weight_slot);
    if (n <= 0 || n_pts <= 0 || !feat || !pooled_out || (feat_dtype != ZS_BF16 && feat_dtype != ZS_BF16_SPLIT))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_pool_debug arguments");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (feat_dtype == ZS_BF16_SPLIT)
        return zs_score_tc3(ctx, weight_slot, feat, n, n_pts, pooled_out, h1_out, h2_out, (cudaStream_t)stream);
    return launch_tc<false>(ctx, weight_slot, (const __nv_bfloat16*)feat, n, n_pts, pooled_out, h1_out, h2_out, fused_args{},
                            (cudaStream_t)stream);
}
