// prefilter_tc.cuh -- tensor-core prefilter (tcgen05.mma kind::f8f6f4, sm_100a).
//
// Same contract as prefilter_kernel (kernels.cuh): emit a SUPERSET of the windows the reference
// accepts (cscore.c:340-389); the fp64 exact stage decides.  Formulation:
//
//     D[128 windows x 256 motif-strands] = A[128 x K] * B[256 x K]^T,   K = 4 * Lpad  (e4m3 x e4m3 -> f16)
//
//   A  the one-hot sequence, never materialised as a matrix: shared memory holds a stream of
//      4 bytes per base (e4m3 1.0 in the byte of the base's code, all zero for a non-ACGT base);
//      row r of A is the stream at byte offset 16 r, the window starting 4 bases after row r-1's.
//      In the no-swizzle K-major canonical layout ((8,m),2):((16 B,SBO),LBO) that is SBO = 128 B,
//      LBO = 16 B -- overlapping rows behind one descriptor.  Four copies of the stream shifted
//      by 0..3 bases give the windows at every start.
//   B  per (motif, strand) column the quantised "budget minus deficit" entries
//          e[b][c] = RU_e4m3( beta - s * (colmax_c - V[b][c]) ),   sum_c beta = s * (colmax_sum - T) + margin
//      so that  raw >= T  implies  sum_c e[b_c][c] >= margin > 0: the candidate test is the
//      accumulator's sign bit.  The products are exact (1.0 x e4m3); each K = 32 step adds its
//      exact partial sum to the f16 accumulator with one round-to-nearest (measured,
//      profiles/r1_tc_toeplitz.txt).  For a window with sum >= 0 every partial sum lies in
//      [-C, C], C = L * beta <= 512, so the <= 4 roundings cost <= 4 * 0.125 and the margin C/64
//      >= 5 keeps the sign.  f16 accumulators are read back two per register (.pack::16b), which
//      halves the epilogue's ALU work; the TMEM read itself costs the same as f32.
//
// One CTA per SM, warp-specialised: warp 0 issues the MMAs, warps 1-3 expand packed codes into
// the one-hot streams (3-slot ring), warps 4-11 read the accumulators back from TMEM (two
// 256-column buffers), AND-reduce the sign bits and emit candidates on the rare set flag.
#pragma once
#include "common.cuh"

namespace msb {

constexpr int kTcTileBases = 512;                       // window starts per position tile
constexpr int kTcStreamBases = kTcTileBases + 32;       // + halo for the longest motif
constexpr int kTcStreamBytes = kTcStreamBases * 4;      // 2176 B per shifted copy
constexpr int kTcSlotBytes = 4 * kTcStreamBytes;        // 8704 B
constexpr int kTcSlots = 3;
constexpr int kTcIgnoreOff = kTcSlots * kTcSlotBytes;   // 26112: 3 x 16 ignore words
constexpr int kTcBOff = 27 * 1024;                      // B tiles start here (1024-aligned)
constexpr int kTcUnitBytes = 8192;                      // one K=32 step of a 256-column tile
constexpr int kTcMaxUnits = 22;                         // 8 KB K-steps of B per batch (180 KB of shared memory)
constexpr int kTcMaxTiles = kTcMaxUnits;                // tiles per batch
constexpr int kTcStageOff = kTcBOff + kTcMaxUnits * kTcUnitBytes;   // candidate staging, per epilogue warp
constexpr int kTcStageCap = 256;                        // staged candidate keys per warp (8 B each)
constexpr int kTcCols = 256;                            // motif-strand columns per tile
constexpr int kTcThreads = 384;
constexpr int kTcEpiWarps = 8;
constexpr int kTcSmemBytes = kTcStageOff + kTcEpiWarps * kTcStageCap * 8;   // dynamic shared memory of the kernel
constexpr int kTcStaticSmemReserve = 1024;             // static __shared__ (barriers) + alignment slack

struct TcBatch {
    uint32_t n_tiles;
    uint32_t b_off;                  // byte offset of the batch's tiles in the global B buffer
    uint32_t b_bytes;
    uint32_t first_tile;             // global tile index of the batch's first tile (col_info row)
    uint8_t ks[kTcMaxTiles];         // K steps (of 32 = 8 bases) per tile
    uint8_t unit_off[kTcMaxTiles];   // tile start within the batch, in 8 KB units
};

struct TcParams {
    SeqView seq;
    const uint8_t *btab;       // all tiles: [ks][kchunk16 (2)][column (256)][16 B]
    TcBatch batch;
    int32_t lmax_all;
    int32_t emit_dirty;
    int32_t any_zero_hit;
    uint64_t *cand;            // raw candidates: key(tile * 256 + column, packed position, 0)
    int64_t cand_cap;
    int64_t *dirty;
    int64_t dirty_cap;
    unsigned long long *counters;
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major, no swizzle (version 1 descriptor): start address, K-chunk byte offset, 8-row-group byte offset
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((addr & 0x3FFFFu) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

#define MSB_TC_LD32(r, addr)                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                          \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                          \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(addr))

#define MSB_TC_LD32P(r, addr)                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "                                \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                          \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(addr))

__device__ __forceinline__ uint32_t and32(const uint32_t (&r)[32]) {
    uint32_t a = r[0] & r[1];
#pragma unroll
    for (int j = 2; j < 32; j += 2) a &= r[j] & r[j + 1];
    return a;
}

// r[j] packs the f16 accumulators of columns 2j (low half) and 2j+1 (high half).
// Bit j of the result = column 2j has a clear sign bit, bit 32 + j = column 2j+1 has.
__device__ __forceinline__ uint64_t sign_clear_mask(const uint32_t (&r)[32]) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint32_t n = ~r[j];
        lo |= ((n >> 15) & 1u) << j;
        hi |= (n >> 31) << j;
    }
    return ((uint64_t) hi << 32) | lo;
}

// Per-warp candidate staging in shared memory.  The epilogue must never wait on global memory
// (a warp that stalls keeps the TMEM buffer from being recycled), so candidate keys are appended
// to a shared-memory buffer with warp shuffles only and flushed with one global atomic per
// kTcStageCap keys.  Column -> (motif, strand) decoding and the sequence-end check happen in the
// exact stage.  All 32 lanes call these.
struct Stage {
    uint64_t *buf;        // this warp's kTcStageCap keys in shared memory
    uint32_t cnt;         // warp-uniform
};

__device__ __forceinline__ void stage_flush(const TcParams &P, Stage &st, int lane) {
    if (st.cnt == 0) return;
    __syncwarp();
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.counters + 0, (unsigned long long) st.cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t i = lane; i < st.cnt; i += 32)
        if ((int64_t) (base + i) < P.cand_cap) P.cand[base + i] = st.buf[i];
    __syncwarp();
    st.cnt = 0;
}

// Bit j of m (this lane): column colid0 + 2j of window p is a candidate; bit 32 + j: column
// colid0 + 2j + 1 (the layout sign_clear_mask produces for one packed 64-column load).
__device__ __forceinline__ uint32_t mask_col(int b) { return b < 32 ? 2u * b : 2u * (b - 32) + 1u; }

__device__ __noinline__ void stage_masks(const TcParams &P, Stage &st, uint64_t m, uint32_t colid0, int64_t p, int lane) {
    const uint32_t n = __popcll(m);
    const uint32_t total = __reduce_add_sync(0xffffffffu, n);
    if (total == 0) return;
    if (total > 64) {
        // dense hits (cutoffs that admit most windows): straight to global memory
        if (n) {
            unsigned long long at = atomicAdd(P.counters + 0, (unsigned long long) n);
            while (m) {
                const int b = __ffsll((long long) m) - 1;
                m &= m - 1;
                if ((int64_t) at < P.cand_cap) P.cand[at] = make_key(colid0 + mask_col(b), p, 0);
                at++;
            }
        }
        return;
    }
    if (st.cnt + total > kTcStageCap) stage_flush(P, st, lane);
    uint32_t my_off = 0, running = 0;
    unsigned ball = __ballot_sync(0xffffffffu, n > 0);
    while (ball) {
        const int src = __ffs(ball) - 1;
        ball &= ball - 1;
        const uint32_t nn = __shfl_sync(0xffffffffu, n, src);
        if (lane == src) my_off = running;
        running += nn;
    }
    uint32_t at = st.cnt + my_off;
    while (m) {
        const int b = __ffsll((long long) m) - 1;
        m &= m - 1;
        st.buf[at++] = make_key(colid0 + mask_col(b), p, 0);
    }
    st.cnt += total;
}

}  // namespace tc

__global__ void __launch_bounds__(kTcThreads, 1)
prefilter_tc_kernel(const __grid_constant__ TcParams P) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_stream = smem;
    uint32_t *s_ignore = reinterpret_cast<uint32_t *>(smem + kTcIgnoreOff);
    uint8_t *s_b = smem + kTcBOff;
    __shared__ uint64_t bar_stream_full[kTcSlots], bar_stream_empty[kTcSlots], bar_tmem_full[2], bar_tmem_empty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ uint32_t s_skip[kTcSlots];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SeqView &S = P.seq;
    const int64_t n_ptiles = (S.total_packed + kTcTileBases - 1) / kTcTileBases;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTcSlots; i++) { mbar_init(&bar_stream_full[i], 3); mbar_init(&bar_stream_empty[i], 1 + kTcEpiWarps); }
        for (int i = 0; i < 2; i++) { mbar_init(&bar_tmem_full[i], 1); mbar_init(&bar_tmem_empty[i], kTcEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // the batch's B tiles -> shared memory, then visible to the async (tensor core) proxy
        const uint4 *src = reinterpret_cast<const uint4 *>(P.btab + P.batch.b_off);
        uint4 *dst = reinterpret_cast<uint4 *>(s_b);
        const uint32_t n16 = P.batch.b_bytes >> 4;
        for (uint32_t i = threadIdx.x; i < n16; i += kTcThreads) dst[i] = __ldg(src + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem_base;
    const uint32_t NT = P.batch.n_tiles;

    if (warp == 0) {
        // ---- MMA issuer -------------------------------------------------------------------------
        if (lane == 0) {
            // D = F16 (0 << 4), A = B = E4M3 (0), both K-major, N = 256 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
            const uint32_t idesc = ((uint32_t) (kTcCols >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
            const uint32_t b_base = smem_u32(s_b);
            uint32_t u = 0, it = 0;
            for (int64_t t = blockIdx.x; t < n_ptiles; t += gridDim.x, it++) {
                const uint32_t slot = it % kTcSlots;
                mbar_wait(&bar_stream_full[slot], (it / kTcSlots) & 1);
                fence_after();
                if (!s_skip[slot]) {
                    const uint32_t a_base = smem_u32(s_stream + slot * kTcSlotBytes);
                    for (uint32_t s = 0; s < 4; s++) {
                        for (uint32_t nt = 0; nt < NT; nt++, u++) {
                            const uint32_t buf = u & 1;
                            mbar_wait(&bar_tmem_empty[buf], ((u >> 1) & 1) ^ 1);
                            fence_after();
                            const uint32_t ks_n = P.batch.ks[nt];
                            const uint32_t bt = b_base + (uint32_t) P.batch.unit_off[nt] * kTcUnitBytes;
                            for (uint32_t ks = 0; ks < ks_n; ks++) {
                                const uint64_t ad = make_desc(a_base + s * kTcStreamBytes + ks * 32, 16, 128);
                                const uint64_t bd = make_desc(bt + ks * kTcUnitBytes, 4096, 128);
                                umma_f8(tmem + buf * kTcCols, ad, bd, idesc, ks > 0);
                            }
                            umma_commit(&bar_tmem_full[buf]);
                        }
                    }
                }
                umma_commit(&bar_stream_empty[slot]);   // fires when the MMAs that read the slot are done
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ---- producers: packed codes -> 4 shifted one-hot streams, ignore bits, dirty windows -----
        const int tid_p = (warp - 1) * 32 + lane;
        const uint32_t horizon = P.lmax_all >= 32 ? 0xffffffffu : ((1u << P.lmax_all) - 1u);
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < n_ptiles; t += gridDim.x, it++) {
            const uint32_t slot = it % kTcSlots;
            mbar_wait(&bar_stream_empty[slot], ((it / kTcSlots) & 1) ^ 1);
            const int64_t tile_start = t * kTcTileBases;
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_stream + slot * kTcSlotBytes);
            for (int a = tid_p; a < kTcStreamBases + 3; a += 96) {
                const int64_t p = tile_start + a;
                uint32_t word = 0;
                if (p < S.total_packed) {
                    const uint32_t code = (__ldg(S.codes + (p >> 4)) >> ((p & 15) * 2)) & 3u;
                    const uint32_t isn = (__ldg(S.nmask + (p >> 5)) >> (p & 31)) & 1u;
                    word = isn ? 0u : (0x38u << (8 * code));   // e4m3 1.0 in the byte of the base
                }
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const int i = a - s;
                    if (i >= 0 && i < kTcStreamBases) dst[s * kTcStreamBases + i] = word;
                }
            }
            if (warp == 1) {
                uint32_t ign = 0xffffffffu;
                if (lane < 16) {
                    const int64_t q0 = tile_start + lane * 32;
                    if (q0 < S.total_packed) {
                        const int64_t blk = q0 >> 5;
                        const int64_t s = __ldg(S.blk_seq + blk);
                        const int64_t j0 = q0 - __ldg(S.poff + s);
                        const int64_t left = (int64_t) __ldg(S.len + s) - j0;
                        const int nvalid = left <= 0 ? 0 : (left >= 32 ? 32 : (int) left);
                        const uint32_t valid = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
                        const uint64_t m64 = ((uint64_t) __ldg(S.nmask + blk + 1) << 32) | __ldg(S.nmask + blk);
                        uint32_t dirtym = 0, emitm = 0;
#pragma unroll
                        for (int k = 0; k < 32; k++) {
                            const uint32_t bits = (uint32_t) (m64 >> k) & horizon;
                            dirtym |= (bits != 0 ? 1u : 0u) << k;
                            emitm |= ((bits != 0 && !(bits == horizon && !P.any_zero_hit)) ? 1u : 0u) << k;
                        }
                        ign = dirtym | ~valid;
                        if (P.emit_dirty) {
                            // windows that touch a non-ACGT base go to the exact kernel directly
                            uint32_t e = emitm & valid;
                            while (e) {
                                const int k = __ffs(e) - 1;
                                e &= e - 1;
                                const unsigned long long sl = atomicAdd(P.counters + 1, 1ull);
                                if ((int64_t) sl < P.dirty_cap) P.dirty[sl] = q0 + k;
                            }
                        }
                    }
                    s_ignore[slot * 16 + lane] = ign;
                }
                const bool all_ign = __all_sync(0xffffffffu, ign == 0xffffffffu);
                if (lane == 0) s_skip[slot] = all_ign ? 1u : 0u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_stream_full[slot]);
        }
    } else {
        // ---- epilogue: warp reads TMEM lanes 32 q .. 32 q + 31 (its 32 windows), 128 of the 256 columns
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        Stage stg;
        stg.buf = reinterpret_cast<uint64_t *>(smem + kTcStageOff) + (warp - 4) * kTcStageCap;
        stg.cnt = 0;
        uint32_t u = 0, it = 0;
        for (int64_t t = blockIdx.x; t < n_ptiles; t += gridDim.x, it++) {
            const uint32_t slot = it % kTcSlots;
            mbar_wait(&bar_stream_full[slot], (it / kTcSlots) & 1);
            const uint32_t skip = s_skip[slot];
            const uint32_t ign4 = (s_ignore[slot * 16 + (row >> 3)] >> ((row & 7) * 4)) & 0xFu;
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_stream_empty[slot]);
            if (skip) continue;
            const int64_t tile_start = t * kTcTileBases;
            for (uint32_t s = 0; s < 4; s++) {
                const bool live = !((ign4 >> s) & 1u);
                const int64_t p = tile_start + 4 * row + s;
                for (uint32_t nt = 0; nt < NT; nt++, u++) {
                    const uint32_t buf = u & 1;
                    mbar_wait(&bar_tmem_full[buf], (u >> 1) & 1);
                    fence_after();
                    const uint32_t taddr = tmem + buf * kTcCols + h * 128 + ((uint32_t) (q * 32) << 16);
                    const uint32_t colid = (P.batch.first_tile + nt) * kTcCols + h * 128;
                    uint32_t r0[32], r1[32];
                    MSB_TC_LD32P(r0, taddr);        // columns h*128 + [0, 64), two per register
                    MSB_TC_LD32P(r1, taddr + 64);   // columns h*128 + [64, 128)
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tmem_empty[buf]);   // accumulators are in registers
                    const bool f0 = live && (and32(r0) & 0x80008000u) != 0x80008000u;
                    const bool f1 = live && (and32(r1) & 0x80008000u) != 0x80008000u;
                    if (__any_sync(0xffffffffu, f0 || f1)) {
                        if (__any_sync(0xffffffffu, f0)) stage_masks(P, stg, f0 ? sign_clear_mask(r0) : 0ull, colid, p, lane);
                        if (__any_sync(0xffffffffu, f1)) stage_masks(P, stg, f1 ? sign_clear_mask(r1) : 0ull, colid + 64, p, lane);
                    }
                }
            }
        }
        stage_flush(P, stg, lane);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace msb
