// Kernel 2 (fp32-accurate): shared per-point MLP 8 -> 64 -> 128 -> 1024 + max over points on tcgen05 with every
// operand split into two bf16 terms, x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits), and every
// product computed as  hi.hi + lo.hi + hi.lo  with fp32 accumulation in TMEM (the lo.lo term, 2^-16 relative, is
// dropped).  Measured against the fp32 oracle the scores agree to ~1.2e-5 of max|score| (budget 1e-4): this is the
// 1e-4 path of the scorer, `model({"point_x": ...})`, python/ossid/utils/zephyr_utils.py:34, and the re-rank that
// makes FrameScorer's top-1 independent of the bf16 rounding of the fast path.
//
// Same machine as zs_score_tc.cu (persistent CTA pairs, cta_group::2 MMAs with M = 256, warp-specialised roles,
// points on TMEM lanes for layers 1-2, channels on lanes for layer 3, no padding rows) with three differences:
//   * W3 as hi + lo is 512 KB, twice what a pair's shared memory holds next to two H2 buffers, so a pair keeps ONE
//     QUARTER of the output channels (256: 128 per CTA, hi + lo = 64 KB per CTA) resident and four pairs share a
//     hypothesis; layers 1-2 (6 % of the work) are recomputed by each of the four.  Pair p -> quarter p & 3,
//     hypotheses (p >> 2) + j * (pairs / 4).
//   * features arrive as two bf16 planes per hypothesis, [n][2][N][8] (ZS_BF16_SPLIT): the hi plane is K columns 0-7
//     of the layer-1 operand and the lo plane K columns 8-15 (the K padding of the bf16 kernel), so layer 1 is
//     [Xhi|Xlo].[W1hi|W1hi]^T + [Xhi|Xlo].[W1lo|0]^T: two MMAs instead of three.
//   * the front epilogues write every activation as two operand tiles (hi, lo); all eight epilogue warps share that
//     work (columns split between warps 0-3 and 4-7) and warps 4-7 also drain the two layer-3 accumulators of a
//     pair-tile, because here the L1 -> H1 -> L2 -> H2 chain, not layer 3, is the critical path.
// MMAs per pair-tile (256 points): 2 (N=64) + 3*4 + 3*16 (N=128) for 256 channels, i.e. 3.2x the bf16 kernel's tensor
// work per hypothesis.
#include <cuda.h>

#include "zs_common.cuh"
#define EXP(bit) false
#include "zs_tc.cuh"

namespace {

using namespace zs_tc;

constexpr int kTile = 128;               // points per CTA per pair-tile
constexpr int kPairTile = 2 * kTile;
constexpr int kThreads3 = 320;
constexpr int kStages = 3;

// ---- shared-memory map (bytes from a 1024-aligned base) ---------------------------------------
constexpr uint32_t kSmW3 = 0;                          // [term hi, lo][k-half] x 16 KB: this CTA's 128 channels of W3
constexpr uint32_t kSmW2 = kSmW3 + 65536;              // [term] x 8 KB: this CTA's 64 rows of W2
constexpr uint32_t kSmA3 = kSmW2 + 16384;              // 2 buffers x ([term][k-half] x 16 KB): H2; H1 hi / lo alias +0 / +32 KB
constexpr uint32_t kA3Buf = 65536, kA3Term = 32768;
constexpr uint32_t kSmW1 = kSmA3 + 2 * kA3Buf;         // [W1hi|W1hi] 1 KB, [W1lo|0] 1 KB: this CTA's 32 rows of W1
constexpr uint32_t kSmX = kSmW1 + 2048;                // 3 stages x (hi plane 2 KB | lo plane 2 KB)
constexpr uint32_t kSmB1 = kSmX + kStages * 4096;      // 64 floats
constexpr uint32_t kSmB2 = kSmB1 + 256;                // 128 floats
constexpr uint32_t kSmBar = kSmB2 + 512;               // mbarriers
constexpr uint32_t kSmTmemPtr = kSmBar + 256;
constexpr uint32_t kSmBytes = kSmTmemPtr + 16;
constexpr uint32_t kSmAlloc = kSmBytes + 1024;         // slack for manual 1024-B alignment
static_assert(kSmAlloc <= 232448, "exceeds 227 KB of shared memory per CTA");

enum : int {
    BAR_W_FULL = 0, BAR_X_FULL = 1 /*3*/, BAR_X_EMPTY = 4 /*3*/, BAR_D1_FULL = 7, BAR_A2_FULL = 8, BAR_D2_FULL = 9,
    BAR_A3_FULL = 10 /*2*/, BAR_A3_EMPTY = 12 /*2*/, BAR_D3_FULL = 14 /*3*/, BAR_D3_EMPTY = 17 /*3*/,
    BAR_XP_FULL = 20 /*3*/, BAR_WP_FULL = 23, BAR_COUNT = 24
};

// D1 has its own columns here (the bf16 kernel aliases it into D2): layer 1 of pair-tile i+1 is issued right behind
// layer 2 of pair-tile i, so its accumulator is waiting when the epilogue warps come back from H2(i) - one commit ->
// wake-up latency less in the L1 -> H1 -> L2 -> H2 chain that bounds this kernel.  Two layer-3 accumulators remain.
constexpr uint32_t kColD2 = 0, kColD1 = 128, kColD3 = 192;
constexpr int kD3Bufs = 2;
constexpr uint32_t kTmemCols = 512;

// global image of the split operands (bytes): W3 [quarter][rank][term][k-half] x 16 KB, W2 [rank][term] x 8 KB,
// W1 [rank][{hi|hi, lo|0}] x 1 KB
constexpr size_t kImg3W2 = 4 * 2 * 4 * 16384, kImg3W1 = kImg3W2 + 2 * 2 * 8192, kImg3Bytes = kImg3W1 + 2 * 2 * 1024;

// two accumulator columns -> relu(acc + bias) split into a bf16x2 of high parts and a bf16x2 of the remainders
// (FADD2 + 2 FMNMX + F2FP, then the remainder as one FFMA2 against the unpacked high parts + F2FP)
__device__ __forceinline__ void split_pack(uint32_t a0, uint32_t a1, float2 b, uint32_t& hi, uint32_t& lo) {
    uint64_t acc, bias, sum, hf, vr, l2, m1;
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(bias) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum) : "l"(acc), "l"(bias));
    float v0, v1, l0, l1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(v0), "=f"(v1) : "l"(sum));
    v0 = fmaxf(v0, 0.f);
    v1 = fmaxf(v1, 0.f);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    asm("mov.b64 %0, {%1,%2};" : "=l"(hf) : "r"(hi << 16), "r"(hi & 0xffff0000u));
    asm("mov.b64 %0, {%1,%2};" : "=l"(vr) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(m1) : "f"(-1.f), "f"(-1.f));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(l2) : "l"(hf), "l"(m1), "l"(vr));      // v - hi, exact
    asm("mov.b64 {%0,%1}, %2;" : "=f"(l0), "=f"(l1) : "l"(l2));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}

// 32 accumulator columns of one point -> 4 swizzled 16-byte chunks in the hi tile and 4 in the lo tile
template <bool kDebug>
__device__ __forceinline__ void epi_store32_split(const uint32_t (&v)[32], const float* __restrict__ bias, uint8_t* tile_hi,
                                                  uint8_t* tile_lo, uint32_t r, uint32_t chunk0, float* __restrict__ dbg) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        const float4 bA = *reinterpret_cast<const float4*>(bias + c8 * 8);
        const float4 bB = *reinterpret_cast<const float4*>(bias + c8 * 8 + 4);
        uint4 h, l;
        split_pack(v[c8 * 8 + 0], v[c8 * 8 + 1], make_float2(bA.x, bA.y), h.x, l.x);
        split_pack(v[c8 * 8 + 2], v[c8 * 8 + 3], make_float2(bA.z, bA.w), h.y, l.y);
        split_pack(v[c8 * 8 + 4], v[c8 * 8 + 5], make_float2(bB.x, bB.y), h.z, l.z);
        split_pack(v[c8 * 8 + 6], v[c8 * 8 + 7], make_float2(bB.z, bB.w), h.w, l.w);
        const uint32_t off = sw128_off(r, chunk0 + c8);
        *reinterpret_cast<uint4*>(tile_hi + off) = h;
        *reinterpret_cast<uint4*>(tile_lo + off) = l;
        if (kDebug && dbg) {                     // what the next layer reads: hi + lo
            const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dbg[c8 * 8 + 2 * e] = __uint_as_float(hh[e] << 16) + __uint_as_float(ll[e] << 16);
                dbg[c8 * 8 + 2 * e + 1] = __uint_as_float(hh[e] & 0xffff0000u) + __uint_as_float(ll[e] & 0xffff0000u);
            }
        }
    }
}

template <bool kDebug>
__global__ void __launch_bounds__(kThreads3, 1)
zs_k_mlp_tc3(const uint8_t* __restrict__ feat, int n, int N, const uint8_t* __restrict__ wimg,
             const float* __restrict__ wf32, float* __restrict__ pooled, float* __restrict__ dbg_h1,
             float* __restrict__ dbg_h2, const int32_t* __restrict__ n_dev, int n_off) {
    n = zs_dyn_count(n_dev, n_off, n);          // zs_set_dynamic_count: the count may live on the device
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - smem_u32(smem_raw));
    auto bar = [&](int i) { return sbase + kSmBar + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_rank();                 // 0 = leader (issues every MMA of the pair)
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int quarter = pair & 3, group = pair >> 2, n_groups = n_pairs >> 2;
    const int T = (N + kPairTile - 1) / kPairTile;        // pair-tiles per hypothesis
    const int n_loc = group < n ? (n - group + n_groups - 1) / n_groups : 0;
    const int total = n_loc * T;
    const size_t hyp_bytes = (size_t)N * 32;              // one hypothesis: hi plane N x 16 B, lo plane N x 16 B

    // ---- one-time setup ------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(bar(BAR_W_FULL), 1);
        mbar_init(bar(BAR_WP_FULL), 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar(BAR_X_FULL + s), 1); mbar_init(bar(BAR_X_EMPTY + s), 1); mbar_init(bar(BAR_XP_FULL + s), 1);
        }
        mbar_init(bar(BAR_D1_FULL), 1); mbar_init(bar(BAR_D2_FULL), 1);
        mbar_init(bar(BAR_A2_FULL), 16);                                  // one arrival per epilogue warp of the pair (8 + 8)
        for (int b = 0; b < 2; ++b) { mbar_init(bar(BAR_A3_FULL + b), 16); mbar_init(bar(BAR_A3_EMPTY + b), 1); }
        for (int b = 0; b < kD3Bufs; ++b) { mbar_init(bar(BAR_D3_FULL + b), 1); mbar_init(bar(BAR_D3_EMPTY + b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < (int)(kStages * 4096 / 16); i += kThreads3)     // stale bytes must be finite
        reinterpret_cast<uint4*>(sm + kSmX)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 64; i += kThreads3) reinterpret_cast<float*>(sm + kSmB1)[i] = wf32[ZS_OFF_B1 + i];
    for (int i = tid; i < 128; i += kThreads3) reinterpret_cast<float*>(sm + kSmB2)[i] = wf32[ZS_OFF_B2 + i];
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

    // first point (within its hypothesis) of this CTA's half of pair-tile `tt`: a half-tile that would run past the
    // last point is shifted back to end at it (max-pooling ignores duplicates); N < 128 repeats the hypothesis
    auto tile_p0 = [&](int tt) {
        const int s0 = tt * kPairTile + rank * kTile;
        return s0 + kTile <= N ? s0 : (N >= kTile ? N - kTile : 0);
    };
    auto hyp_of = [&](int j) { return group + j * n_groups; };

    if (warp == 8) {
        // ===== bulk-copy producer =================================================================
        if (lane == 0) {
            mbar_expect_tx(bar(BAR_W_FULL), 65536 + 16384 + 2048);
            const uint8_t* w3 = wimg + (size_t)(quarter * 2 + rank) * 65536;
            for (int c = 0; c < 4; ++c) bulk_g2s(sbase + kSmW3 + c * 16384, w3 + (size_t)c * 16384, 16384, bar(BAR_W_FULL));
            bulk_g2s(sbase + kSmW2, wimg + kImg3W2 + (size_t)rank * 16384, 16384, bar(BAR_W_FULL));
            bulk_g2s(sbase + kSmW1, wimg + kImg3W1 + (size_t)rank * 2048, 2048, bar(BAR_W_FULL));
            // the leader's issuer must also know that the PEER's operands have landed (relayed one tile behind)
            const bool relay = rank != 0;
            const uint32_t wp = leader_addr(bar(BAR_WP_FULL)), xp = leader_addr(bar(BAR_XP_FULL));
            if (relay) { mbar_wait(bar(BAR_W_FULL), 0); mbar_arrive_leader(wp); }
            for (int i = 0; i < total; ++i) {
                const int s = i % kStages, j = i / T, tt = i - j * T;
                const uint8_t* src = feat + (size_t)hyp_of(j) * hyp_bytes + (size_t)tile_p0(tt) * 16;
                const uint32_t dst = sbase + kSmX + s * 4096;
                mbar_wait(bar(BAR_X_EMPTY + s), ((i / kStages) & 1) ^ 1);
                mbar_expect_tx(bar(BAR_X_FULL + s), 4096);
                if (N >= kTile) {
                    bulk_g2s(dst, src, 2048, bar(BAR_X_FULL + s));
                    bulk_g2s(dst + 2048, src + (size_t)N * 16, 2048, bar(BAR_X_FULL + s));
                } else {
                    for (int r = 0; r < kTile; r += N) {    // the whole (short) hypothesis, repeated, both planes
                        const uint32_t bytes = (uint32_t)min(N, kTile - r) * 16;
                        bulk_g2s(dst + r * 16, src, bytes, bar(BAR_X_FULL + s));
                        bulk_g2s(dst + 2048 + r * 16, src + (size_t)N * 16, bytes, bar(BAR_X_FULL + s));
                    }
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
            constexpr uint32_t hi_sw = desc_hi(1024, kLayoutSw128), hi_x = desc_hi(128, kLayoutNone);
            const uint32_t w3_lo = desc_lo(sbase + kSmW3, 16), w2_lo = desc_lo(sbase + kSmW2, 16);
            const uint32_t w1_lo = desc_lo(sbase + kSmW1, 512);
            // layer 3 of pair-tile `it`, point half hh: W3hi.H2hi + W3hi.H2lo + W3lo.H2hi, 8 K-steps each
            auto issue_l3 = [&](int it, int hh) {
                const int q = it * 2 + hh, b = q % kD3Bufs, buf = it & 1;
                const uint32_t a3_lo = desc_lo(sbase + kSmA3 + buf * kA3Buf + hh * 8192, 16), d3 = tmem + kColD3 + b * 128;
                mbar_wait(bar(BAR_D3_EMPTY + b), ((q / kD3Bufs) & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t tw = term == 2 ? 1u : 0u, ta = term == 1 ? 1u : 0u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t kb = k >> 2, kk = k & 3;
                        tc_mma(d3, desc_at(w3_lo, hi_sw, (tw * 2 + kb) * 16384 + kk * 32),
                               desc_at(a3_lo, hi_sw, ta * kA3Term + kb * 16384 + kk * 32), idesc_128, (term | k) != 0);
                    }
                }
                tc_commit(bar(BAR_D3_FULL + b));
                if (hh == 1) tc_commit(bar(BAR_A3_EMPTY + buf));
            };
            auto issue_l1 = [&](int it) {                      // [Xhi|Xlo].[W1hi|W1hi]^T + [Xhi|Xlo].[W1lo|0]^T
                const int s = it % kStages;
                mbar_wait(bar(BAR_X_FULL + s), (it / kStages) & 1);
                mbar_wait(bar(BAR_XP_FULL + s), (it / kStages) & 1);
                tc_fence_after();
                const uint64_t xd = desc_at(desc_lo(sbase + kSmX + s * 4096, 2048), hi_x, 0);   // K 0-7 = hi plane, 8-15 = lo plane
                tc_mma(tmem + kColD1, xd, desc_at(w1_lo, hi_x, 0), idesc_l1, 0);
                tc_mma(tmem + kColD1, xd, desc_at(w1_lo, hi_x, 1024), idesc_l1, 1);
                tc_commit(bar(BAR_X_EMPTY + s));
                tc_commit(bar(BAR_D1_FULL));
            };
            if (total > 0) issue_l1(0);
            for (int i = 0; i <= total; ++i) {
                // H2 of pair-tile i-1 is ready in both CTAs and both epilogues 2 have drained D2
                if (i >= 1) { mbar_wait(bar(BAR_A3_FULL + ((i - 1) & 1)), ((i - 1) >> 1) & 1); issue_l3(i - 1, 0); }
                if (i < total) {
                    const uint32_t a2_lo = desc_lo(sbase + kSmA3 + (i & 1) * kA3Buf, 16);
                    mbar_wait(bar(BAR_A2_FULL), i & 1);            // H1(i) is in place, so D1 has been read
                    tc_fence_after();
#pragma unroll
                    for (int term = 0; term < 3; ++term) {       // H1hi.W2hi + H1lo.W2hi + H1hi.W2lo
                        const uint32_t ta = term == 1 ? 1u : 0u, tw = term == 2 ? 1u : 0u;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tc_mma(tmem + kColD2, desc_at(a2_lo, hi_sw, ta * kA3Term + kk * 32),
                                   desc_at(w2_lo, hi_sw, tw * 8192 + kk * 32), idesc_128, (term | kk) != 0);
                    }
                    tc_commit(bar(BAR_D2_FULL));
                    if (i + 1 < total) issue_l1(i + 1);           // early: D1 is its own TMEM range
                }
                if (i >= 1) issue_l3(i - 1, 1);
            }
        }
    } else {
        // ===== epilogue warps ======================================================================
        // Warps 0-3 ("front A") and 4-7 ("front B") own the same four TMEM lane quadrants and split the COLUMNS of the
        // front epilogues: D1 -> H1 channels [0,32) / [32,64), D2 -> H2 channels [0,64) / [64,128) (= the two K halves
        // of layer 3), each written as a hi and a lo operand tile.  Front B also drains the layer-3 accumulators
        // (max over points): in this kernel a pair-tile has only two of them, so those warps have the slack, and
        // halving the per-warp epilogue work is what takes the L1 -> H1 -> L2 -> H2 chain off the critical path.
        // Front B's order per iteration follows the issuer's: H1(i), drain D3(i-1, 0), H2(i), drain D3(i-1, 1).
        const bool front_b = warp >= 4;
        const int wq = warp & 3;                                        // TMEM lane quadrant == warp % 4
        const uint32_t r = (uint32_t)(wq * 32 + lane);                  // TMEM lane = point row (fronts) = channel (max-pool)
        const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
        const float* b1 = reinterpret_cast<const float*>(sm + kSmB1) + (front_b ? 32 : 0);
        const float* b2 = reinterpret_cast<const float*>(sm + kSmB2) + (front_b ? 64 : 0);
        const uint32_t a2_full = leader_addr(bar(BAR_A2_FULL)), a3_full = leader_addr(bar(BAR_A3_FULL));
        const uint32_t d3_empty = leader_addr(bar(BAR_D3_EMPTY));
        const bool dump = kDebug && quarter == 0;                       // the four pairs of a hypothesis compute identical H1 / H2
        const int ch = quarter * 256 + rank * 128 + (int)r;
        const float b3 = __ldg(wf32 + ZS_OFF_B3 + ch);
        float qq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto drain = [&](int it, int hh) {                              // D3[channel lane][point column] -> running max
            const int q = it * 2 + hh, b = q % kD3Bufs;
            mbar_wait(bar(BAR_D3_FULL + b), (q / kD3Bufs) & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32], v2[32], v3[32];                    // all 128 point columns in flight, one wait
            const uint32_t a = lane_addr + kColD3 + b * 128;
            tc_ld32(a, v0); tc_ld32(a + 32, v1); tc_ld32(a + 64, v2); tc_ld32(a + 96, v3);
            tc_wait_ld();
            tc_fence_before();                                          // the accumulator is free once it sits in registers
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(d3_empty + 8u * b);
            max32(v0, qq); max32(v1, qq); max32(v2, qq); max32(v3, qq);
        };
        for (int i = 0; i <= total; ++i) {
            const int buf = i & 1, j = i / T, tt = i - j * T;
            uint8_t* a3 = sm + kSmA3 + buf * kA3Buf;
            long long drow = 0;
            if (kDebug && i < total) drow = (long long)hyp_of(j) * N + tile_p0(tt) + (N >= kTile ? (int)r : (int)r % N);
            if (i < total) {
                mbar_wait(bar(BAR_A3_EMPTY + buf), ((i >> 1) & 1) ^ 1); // layer 3 of pair-tile i-2 has released this buffer
                mbar_wait(bar(BAR_D1_FULL), i & 1);
                tc_fence_after();
                uint32_t v0[32];
                tc_ld32(lane_addr + kColD1 + (front_b ? 32 : 0), v0);
                tc_wait_ld();
                epi_store32_split<kDebug>(v0, b1, a3, a3 + kA3Term, r, front_b ? 4 : 0,
                                          (dump && dbg_h1) ? dbg_h1 + drow * 64 + (front_b ? 32 : 0) : nullptr);
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(a2_full);
            }
            if (front_b && i >= 1) drain(i - 1, 0);
            if (i < total) {
                mbar_wait(bar(BAR_D2_FULL), i & 1);
                tc_fence_after();
                const int kb = front_b ? 1 : 0;                         // 64 channels = one K half of layer 3
                uint32_t v0[32], v1[32];
                tc_ld32(lane_addr + kColD2 + kb * 64, v0);
                tc_ld32(lane_addr + kColD2 + kb * 64 + 32, v1);
                tc_wait_ld();
                float* d = (dump && dbg_h2) ? dbg_h2 + drow * 128 + kb * 64 : nullptr;
                epi_store32_split<kDebug>(v0, b2, a3 + kb * 16384, a3 + kA3Term + kb * 16384, r, 0, d);
                epi_store32_split<kDebug>(v1, b2 + 32, a3 + kb * 16384, a3 + kA3Term + kb * 16384, r, 4, d ? d + 32 : nullptr);
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(a3_full + 8u * buf);
            }
            if (front_b && i >= 1) {
                drain(i - 1, 1);
                const int jp = (i - 1) / T, ttp = (i - 1) - jp * T;
                if (ttp == T - 1) {                                     // last pair-tile of the hypothesis: bias + ReLU after the max
                    const float mm = fmaxf(fmaxf(qq[0], qq[1]), fmaxf(qq[2], qq[3]));
                    pooled[(size_t)hyp_of(jp) * 1024 + ch] = fmaxf(mm + b3, 0.f);
                    qq[0] = qq[1] = qq[2] = qq[3] = -INFINITY;
                }
            }
        }
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

// ---- weight images: every weight as bf16 hi + bf16 lo, laid out exactly as the tiles sit in shared memory ---------
__device__ __forceinline__ void put_split(uint8_t* img, size_t off_hi, size_t off_lo, float w) {
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    *reinterpret_cast<__nv_bfloat16*>(img + off_hi) = h;
    *reinterpret_cast<__nv_bfloat16*>(img + off_lo) = __float2bfloat16_rn(w - __bfloat162float(h));
}

__global__ void zs_k_build_images3(const float* __restrict__ w, uint8_t* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1024 * 128) {                  // W3: channel ch -> quarter, rank, row; K -> k-half, swizzled chunk
        const int ch = i >> 7, k = i & 127;
        const int quarter = ch >> 8, rank = (ch >> 7) & 1, mrow = ch & 127, kb = k >> 6, kc = k & 63;
        const size_t base = (size_t)(quarter * 2 + rank) * 65536 + (size_t)kb * 16384 + sw128_off(mrow, kc >> 3) + (kc & 7) * 2;
        put_split(img, base, base + 32768, w[ZS_OFF_W3 + i]);
    }
    if (i < 128 * 64) {                    // W2: rows 0-63 leader's share of the N side, 64-127 the peer's
        const int ch = i >> 6, k = i & 63;
        const size_t base = kImg3W2 + (size_t)(ch >> 6) * 16384 + sw128_off(ch & 63, k >> 3) + (k & 7) * 2;
        put_split(img, base, base + 8192, w[ZS_OFF_W2 + i]);
    }
    if (i < 64 * 16) {                     // W1: K 0-7 and 8-15 both carry the weight (hi image); lo image has zeros in K 8-15
        const int ch = i >> 4, k = i & 15, chl = ch & 31;
        const size_t core = (size_t)(k >> 3) * 512 + (size_t)(chl >> 3) * 128 + (chl & 7) * 16 + (k & 7) * 2;
        const size_t base = kImg3W1 + (size_t)(ch >> 5) * 2048 + core;
        const float v = w[ZS_OFF_W1 + ch * 8 + (k & 7)];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(img + base) = h;
        *reinterpret_cast<__nv_bfloat16*>(img + base + 1024) = k < 8 ? __float2bfloat16_rn(v - __bfloat162float(h)) : __float2bfloat16_rn(0.f);
    }
}

// fp32 features [n][N][8] -> split planes [n][2][N][8] bf16 (one thread per point: 32 bytes in, 2 x 16 bytes out)
__global__ void zs_k_split_features(const float4* __restrict__ in, long long n_rows, int N, uint4* __restrict__ out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows) return;
    const float4 a = __ldg(in + 2 * g), b = __ldg(in + 2 * g + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = pack2(f[2 * e], f[2 * e + 1]);
        l[e] = pack2(f[2 * e] - __uint_as_float(h[e] << 16), f[2 * e + 1] - __uint_as_float(h[e] & 0xffff0000u));
    }
    const long long hyp = g / N, p = g - hyp * N;
    out[hyp * 2 * N + p] = make_uint4(h[0], h[1], h[2], h[3]);
    out[hyp * 2 * N + N + p] = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace

extern "C" int zs_split_features(zs_ctx* ctx, const float* feat, int n, int n_pts, void* split_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n == 0) return ZS_OK;
    if (n < 0 || n_pts <= 0 || !feat || !split_out || ((uintptr_t)feat & 15) || ((uintptr_t)split_out & 15))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_split_features arguments");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long rows = (long long)n * n_pts;
    zs_k_split_features<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(feat), rows, n_pts, reinterpret_cast<uint4*>(split_out));
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

int zs_tc3_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st) {
    zs_weights& w = ctx->w[slot];
    if (!w.bf16x2 && cudaMalloc(&w.bf16x2, kImg3Bytes) != cudaSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_NOMEM, "split bf16 weight images");
    }
    zs_k_build_images3<<<(1024 * 128 + 255) / 256, 256, 0, st>>>(w.f32, reinterpret_cast<uint8_t*>(w.bf16x2));
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

// feat: ZS_BF16_SPLIT features [n][2][n_pts][8] bf16 -> pooled [n][1024] fp32 (fp32-accurate)
int zs_score_tc3(zs_ctx* ctx, int slot, const void* feat, int n, int n_pts, float* pooled, float* dbg_h1, float* dbg_h2,
                 cudaStream_t st) {
    const zs_weights& w = ctx->w[slot];
    const bool debug = dbg_h1 || dbg_h2;
    auto kernel = debug ? zs_k_mlp_tc3<true> : zs_k_mlp_tc3<false>;
    ZS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmAlloc));
    int groups = (ctx->sm_count / 2) / 4;          // four CTA pairs (one per channel quarter) share a hypothesis
    if (groups > n) groups = n;
    if (groups < 1) groups = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * 8), 1, 1);
    cfg.blockDim = dim3(kThreads3, 1, 1);
    cfg.dynamicSmemBytes = kSmAlloc;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;    // CTA pair = the two SMs of a TPC (cta_group::2 MMAs)
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ZS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, (const uint8_t*)feat, n, n_pts, (const uint8_t*)w.bf16x2,
                                    (const float*)w.f32, pooled, dbg_h1, dbg_h2, ctx->dyn_n, ctx->dyn_off));
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}
